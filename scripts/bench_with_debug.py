import sys
sys.path.insert(0, ".")
from bpmult_b200 import _lib
lib = _lib.load()
dbg = int(sys.argv[1])
lib.bpm_debug_set(0, dbg)
sys.argv = ["bench.py", "--steps", "6", "--warmup", "3"]
import bench
bench.main()
