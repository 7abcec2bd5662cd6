"""CPU oracle for the deephisto_b200 hot path -- TEST INFRASTRUCTURE ONLY.

Plain numpy / pure-Python restatements of the reference algorithms (each function cites the
reference file:line it follows, relative to xubiker/deephisto). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the product
(deephisto_b200/) never does.

Pinning status (see DESIGN.md "Oracle"):
  * dense coordinates, gather + /255, stitch sum map + argmax: PINNED against the unmodified
    reference run through numpy stubs of psimage (oracle/reference_loader.py, tests/golden/).
  * polygon clip area / acceptance, region sampling, coverage sampler draws: PARITY UNPINNED --
    shapely/GEOS is not in the reference tree and the reference RNG is the unseeded global numpy
    generator; these are restatements of the published definitions, cross-checked against an
    independent Sutherland-Hodgman clipper.
"""
