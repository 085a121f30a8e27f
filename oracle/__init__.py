"""CPU oracle for the RefineDet detect / anchor-matching hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker (or the
timed CPU baseline), never as the thing shipped.  The product path
(``refinedet.pytorch_b200``) calls the CUDA C-ABI library and fails loudly when
it is missing.

Parity status: **pinned by reference outputs generated in the authoring
container** (the reference has no tests / golden vectors of its own, SURVEY.md
§4 and §8c).  ``tests/golden/make_golden.py`` imports the unmodified reference
from ``/root/reference`` and records its outputs; ``tests/test_oracle_golden.py``
checks this restatement against those fixtures.
"""
from . import box_oracle  # noqa: F401
