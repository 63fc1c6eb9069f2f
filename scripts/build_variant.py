"""Build an A/B variant of the library: python scripts/build_variant.py NAME -DFOO=1 ...  -> dig_b200/libdig_b200_NAME.so
(select it at run time with DIG_B200_LIB=libdig_b200_NAME.so)."""
import glob, os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, extra = sys.argv[1], sys.argv[2:]
srcs = sorted(glob.glob(os.path.join(ROOT, "dig_b200", "csrc", "*.cu")))
od = os.path.join(ROOT, "build", "variant_" + name)
os.makedirs(od, exist_ok=True)
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"] + extra
objs = [os.path.join(od, os.path.basename(s)[:-3] + ".o") for s in srcs]
def cc(so):
    subprocess.run(["nvcc"] + flags + ["-c", "-o", so[1], so[0]], check=True)
with ThreadPoolExecutor(8) as ex:
    list(ex.map(cc, zip(srcs, objs)))
out = os.path.join(ROOT, "dig_b200", "libdig_b200_%s.so" % name)
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, check=True)
print(out)
