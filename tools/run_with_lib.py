"""Dev helper: run a tool script against another build of the library: python tools/run_with_lib.py <lib.so> <script> [args]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bevgen_b200._native as N
N.LIB_PATH = Path(sys.argv[1]).resolve()
script = sys.argv[2]
sys.argv = [script] + sys.argv[3:]
exec(compile(open(script).read(), script, "exec"))
