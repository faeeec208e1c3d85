"""Builds experimental variants of the engine library beside the default one: csrc/libvar_<name>.so per "name=flags" argument
(tools/gpu/gpu_lib_sweep.py times them; SOLR_B200_LIB selects one).  Example:
  python tools/build_variants.py a_unit="-DUW_GROUP=0" b_qt16="-DGW_QT=16 -DGW_QPER=2" """
import os, sys, subprocess, glob
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import build as b

def one(arg):
    name, flags = arg.split("=", 1)
    out = os.path.join(b.CSRC, "libvar_%s.so" % name)
    cmd = [b.NVCC] + b.NVCC_FLAGS + flags.split() + ["-o", out, os.path.join(b.CSRC, "engine.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, (r.stderr or "")[-2000:]

if __name__ == "__main__":
    if "--clean" in sys.argv:
        for f in glob.glob(os.path.join(b.CSRC, "libvar_*.so")): os.remove(f)
    args = [a for a in sys.argv[1:] if "=" in a]
    with ThreadPoolExecutor(4) as ex:
        for name, rc, err in ex.map(one, args):
            print(name, "rc", rc, err if rc else "")
