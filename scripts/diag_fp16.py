"""Development: accuracy (stage errors vs the oracle) and speed of the camera branch in fp16 storage."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.argv = ["x"]
import importlib.util
spec_ = importlib.util.spec_from_file_location("diag_e2e_mod", os.path.join(os.path.dirname(__file__), "diag_e2e.py"))
src = open(os.path.join(os.path.dirname(__file__), "diag_e2e.py")).read().split('stage_errors("cudnn_tf32")')[0]
ns = {"__name__": "diag", "__file__": os.path.join(os.path.dirname(os.path.abspath(__file__)), "diag_e2e.py")}
exec(compile(src, "diag_e2e_head", "exec"), ns)
ns["stage_errors"]("fp32_tf32", None)
ns["stage_errors"]("fp16", torch.float16)
