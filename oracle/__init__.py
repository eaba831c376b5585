"""CPU oracle for the NPE hot path of mbchang/dynamics.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``dynamics_b200``) never imports this package and
fails loudly when ``libnpe_cuda.so`` is missing.

PARITY UNPINNED: the reference (pure Lua/Torch7) ships no tests, golden vectors
or fixtures for this path, and neither ``th``/LuaJIT nor the Torch7 rocks that
hold the arithmetic (torch7, nn, nngraph, rnn, optim -- all un-vendored and
unpinned, ``README.md:43-52``) exist in this container, so the reference cannot
be executed.  This package is an op-for-op float32 restatement of the Lua call
sites (each function cites the ``file:line`` it follows; upstream-rock semantics
are restated from their published behaviour and flagged as such).  It is
cross-validated internally: the manual backward against torch-CPU autograd and
finite differences, and property tests (masked pair == removed pair, context
permutation invariance, trim identity for C <= 12).
"""
