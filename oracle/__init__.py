"""CPU oracle for the GPEMSR inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  ``gpemsr_b200`` never imports this package.

Parity status:
  * codebook / decoder / blocks / indexer head: PINNED -- ``make_golden.py``
    imports the reference's own ``model/{codebook,decoder,blocks}.py`` in the
    authoring container and the restatement reproduces them bit-for-bit on the
    committed fixtures in ``tests/golden/``.
  * SR tail (``GPEMSR.forward`` lines 441-455): pinned against the reference's
    ``model/GPEMSR.py`` run through ``basicsr_shim`` (``ResidualBlockNoBN`` is
    a third-party BasicSR symbol, restated from its published definition).
  * flow_warp: PARITY UNPINNED at the BasicSR boundary -- BasicSR is an
    un-vendored, un-pinned dependency (PyPI ``basicsr``; upstream v1.4.2
    ``basicsr/archs/arch_util.py``) that is absent from /root/reference and from
    this image.  Its published algorithm is restated in ``flow_warp.py``; the
    ``grid_sample`` half is pinned against ATen (torch CPU) on the fixtures.
"""
