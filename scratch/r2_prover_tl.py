"""ZA_DEBUG_TIMELINE=1 python scratch/r2_prover_tl.py N [log_m]: per-device timeline of za_prover proofs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, za_b200
from za_b200 import synthetic
N = int(sys.argv[1]); log_m = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nc = (1 << log_m) - 2
ni, na, ptr, var, coeff, inputs, aux = synthetic.mul_chain(nc, x0=5); counts = synthetic.pk_counts_for_mul_chain(nc)
pr = za_b200.Prover(list(range(N)))
pr.synthetic_pk(counts["ic"], counts["h"], counts["l"], counts["a"], counts["b_g1"], counts["b_g2"])
pr.set_circuit(ni, na, ptr, var, coeff)
pr.upload_witness(inputs, aux)
for i in range(6):
    t = time.perf_counter()
    p = pr.create_proof(None, None, 11, 13)
    print("proof %d: %.3f ms" % (i, (time.perf_counter() - t) * 1e3), file=sys.stderr, flush=True)
print(pr.info())
