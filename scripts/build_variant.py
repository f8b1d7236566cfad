"""Build a tuning variant of libcpq.so: python scripts/build_variant.py NAME KEY=VAL ...  -> build/variants/libcpq_NAME.so
(run with CPQ_LIB=build/variants/libcpq_NAME.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convopeq_b200 import build
name = sys.argv[1]
defs = dict(kv.split("=") for kv in sys.argv[2:])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print(build.build(force=True, out=os.path.join(root, "build", "variants", f"libcpq_{name}.so"), defines=defs))
