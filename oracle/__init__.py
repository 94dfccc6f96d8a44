"""TEST INFRASTRUCTURE ONLY - not part of the product.

``oracle/`` holds the CPU checker for the NSGP-RePRE hot path:

* ``restated.py``   - torch-CPU / numpy restatement of the reference algorithm
                      (every function cites the reference file:line it follows).
                      This is what travels to the GPU box.
* ``ref_loader.py`` - stub importer that loads the reference's OWN functions
                      read-only from ``/root/reference`` (build container only;
                      that path does not exist on the GPU box).
* ``make_golden.py``- runs the reference's own functions on seeded inputs and
                      writes small fixtures into ``tests/golden/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
package (``nsgp-repre_b200/``) never imports it and has no CPU fallback.

Parity pinning: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4 / 8c).  The restatement is therefore pinned against outputs of
the reference's own functions executed in the build container
(``tests/golden/*.pt`` + ``oracle/make_golden.py``).
"""
