"""CPU oracle for the svgb200 hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Every function here restates, in plain numpy / CPU torch (fp32), one algorithm of the reference
(svg-project/Sparse-VideoGen, mounted at /root/reference in the build container) and cites the
reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product (sparse-videogen_b200/) never does.

Pinning status (see DESIGN.md "Oracle"):
  * attention / masks / placement / permutation / density / dynamic map: checked against the
    reference's own functions imported in the build container (tests/golden/make_golden.py) and
    against committed golden vectors generated from them (tests/golden/*.npz).
  * k-means (batch_kmeans_Euclid), sample_mse: the reference functions need a GPU (Triton) or
    diffusers to import, and the reference has no test for them -> restated from source, parity
    UNPINNED by reference outputs (pinned only by the torch formulation the reference itself keeps
    in comments, kmeans_utils.py:631-635).
"""
