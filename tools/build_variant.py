"""Development aid: builds build/variants/<name>/libtfhe_b200.so with extra -D flags on
blind_rotate.cu / keyswitch.cu (the other objects are reused from cpu-gpu-tfhe_b200/_obj), so
that several compile-time variants can be timed in ONE gpurun call:
    python tools/build_variant.py peel -DTFHE_B200_MAC_PEEL=1
    TFHE_B200_LIB=build/variants/peel/libtfhe_b200.so python tools/quick_bench.py 4736
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpu-gpu-tfhe_b200"))
import build as B

name, flags = sys.argv[1], sys.argv[2:]
B.build(verbose=False)
out = os.path.join(ROOT, "build", "variants", name)
os.makedirs(out, exist_ok=True)
objs = []
for s in B.SOURCES:
    obj = os.path.join(B.OBJ, s[:-3] + ".o")
    if s in ("blind_rotate.cu", "keyswitch.cu", "keyswitch_mma.cu", "engine.cu"):
        obj = os.path.join(out, s[:-3] + ".o")
        cmd = [B.NVCC] + B.FLAGS + flags + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, s), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.exit(r.stdout[-4000:])
        for line in r.stdout.splitlines():
            if "registers" in line or "spill" in line:
                print(s, line.strip())
    objs.append(obj)
lib = os.path.join(out, "libtfhe_b200.so")
subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print("built", lib)
