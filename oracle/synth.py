"""Re-export of the synthetic workload generator (it lives in the product package, ``icrl_b200.synth``, so that
``bench.py``'s GPU arm needs nothing from ``oracle/``; the oracle may import the product, never the reverse)."""
from icrl_b200.synth import *            # noqa: F401,F403
from icrl_b200.synth import H, VOCAB, START, END, word_to_idx, make_weights, a2c_state_dict, make_inputs, make_uniforms  # noqa: F401
