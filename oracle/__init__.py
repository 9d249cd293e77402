"""TEST INFRASTRUCTURE — CPU oracle for the Peritext hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package.  Nothing under ``peritext_b200/`` may import it.
"""
