"""Writes the file-format fixtures with the REFERENCE's own serialisation code (oracle/_ref,
i.e. /root/reference/gpuParallel/tfhe_io.cu compiled as is): a ciphertext file and the text
header of a cloud key.  Run in the build container; the outputs are committed."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import Ref

ref = Ref().keygen((314, 1592, 657))
bits = [1, 0, 1, 1, 0, 0, 1, 0]
c = np.stack([ref.encrypt(b) for b in bits])
v = np.arange(1, len(bits) + 1) * 5.9536e-10
ref.write_ciphertexts(os.path.join(HERE, "ref_ciphertexts.bin"), c, v)
np.savez(os.path.join(HERE, "ref_ciphertexts.npz"), samples=c, variances=v, bits=np.array(bits))
ref.write_cloud_key("/tmp/_golden_cloud.key")
with open("/tmp/_golden_cloud.key", "rb") as f:
    data = f.read(2000)
end = data.index(b"-----END LWEKSPARAMS-----\n") + len(b"-----END LWEKSPARAMS-----\n")
with open(os.path.join(HERE, "ref_cloud_key_header.txt"), "wb") as f:
    f.write(data[:end])
os.remove("/tmp/_golden_cloud.key")
print(data[:end].decode())
