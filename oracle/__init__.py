"""CPU oracle for the A2C caption-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the CPU
baseline being timed.  The product path (``icrl_b200``) never imports it and has
no CPU fallback.

Parity status: the reference (pratikpv/image-captioning-through-rl) ships no tests
and no golden vectors, so this oracle is pinned against the
reference *itself*, executed unmodified in the build container by
``oracle/gen_golden.py`` (committed) which writes ``tests/golden/*.npz``.  The
oracle is checked against those fixtures by ``tests/test_oracle_golden.py`` on
every CPU run, and against the live reference when ``/root/reference`` exists.

Modules
-------
synth        re-export of icrl_b200.synth (seeded synthetic weights / inputs; lives product-side so that
             bench.py's GPU arm imports nothing from oracle/)
ref_port     the reference algorithm *as executed* (per-step prefix re-runs, the
             batch-as-time value/reward RNN calls with carried state) on torch
             CPU library layers -- the CPU baseline that bench.py times
single_pass  the verified single-pass restatement with the cell
             arithmetic written out -- what the CUDA kernels are checked against
"""
