import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icrl_b200 import _lib
torch.cuda.init()
print("fwd max pieces", _lib.call("icrl_chain_tc_max_pieces"), "bwd max pieces", _lib.call("icrl_chain_tc_bwd_max_pieces"))
