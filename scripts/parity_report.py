import sys, json
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from parity_util import run_model_parity
for m in ["LightGCN", "LayerGCN", "FREEDOM", "MGCN", "SMORE"]:
    try:
        print(m, json.dumps(run_model_parity(m, "cuda:0")))
    except Exception as e:
        import traceback; traceback.print_exc()
