"""Build a variant of the library with extra nvcc flags: python examples/bench_scripts/build_variant.py NAME -DFOO ...  -> examples/bench_scripts/NAME.so"""
import os, subprocess, sys, concurrent.futures as cf
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "pime-robust-non-linear-set-point-control-with-reinforcement-learning_b200"))
import build as B
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name)
os.makedirs(out + "_obj", exist_ok=True)
def comp(src):
    o = os.path.join(out + "_obj", src.replace(".cu", ".o"))
    r = subprocess.run([B.nvcc(), *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", o], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return o
with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(comp, B.SOURCES))
r = subprocess.run([B.nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out + ".so", *objs, "-cudart", "static"], capture_output=True, text=True)
assert r.returncode == 0, r.stderr
print(out + ".so")
