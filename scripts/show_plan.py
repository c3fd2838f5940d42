import os, sys
sys.path.insert(0, "/root/repo")
os.environ["IEVM_VERBOSE"] = "1"
import ievm_b200
from ievm_b200 import synthetic as mf
eng = ievm_b200.B200QuantizedResNet.from_converted(mf.static_quantize_fbgemm(mf.make_student()), max_batch=256)
eng.close()
