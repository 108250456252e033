"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's overlap-graph hot path (``oracle/*.c``) and a harness around
the unmodified reference sources (``oracle/_ref``).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this package; the product
package ``alga_b200`` never does.
"""
