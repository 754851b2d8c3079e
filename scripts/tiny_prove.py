"""One prove of a 2^LOG2N-row chain circuit (default 2^6) after a warm-up — target for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from prove_bench import synthetic

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ctx = zk.Context(0)
c = synthetic(log2n)
s = zk.SRS.NewSRS((1 << log2n) + 3, zkp.fr_to_mont([777]), ctx).precompute()
pk = zkp.ProvingKey.SetupRaw(s, log2n, log2n + 2, 1, c["nb_wires"], c["ql"], c["qr"], c["qm"], c["qo"], c["qk"], c["lro"], ctx)
bl = np.zeros(9 * 32, dtype=np.uint8)
bl[::32] = 7
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    pk.Prove(c["sol"], bl)
print("done")
