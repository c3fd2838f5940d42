"""Print the per-layer kernel plan (IEVM_VERBOSE=1) of the INT8 pruned ResNet-18 for a given max_batch.
   python scripts/show_plan.py [max_batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IEVM_VERBOSE"] = "1"
import ievm_b200
from ievm_b200 import synthetic as mf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = ievm_b200.B200QuantizedResNet.from_converted(mf.static_quantize_fbgemm(mf.make_student()), max_batch=n)
eng.close()
