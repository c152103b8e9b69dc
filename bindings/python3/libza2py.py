"""libza2py — the module name of za's Python 3 binding (/root/reference/binding/python3/native/src/lib.rs:9-16), so that
`import libza2py as circom` of /root/reference/binding/python3/test/test.py runs unchanged with this directory on
PYTHONPATH.  Functions: verbose(on) -> bool, setup(circuit_path, pk_path, verifier_type) -> str,
prove(pk_path, inputs_json) -> str, verify(verifying_key_json, proof_with_inputs_json) -> bool; errors are TypeError
with the Debug text of the error, as the reference raises them."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from za_b200.za2c import prove, setup, verbose, verify  # noqa: E402,F401

__doc__ = "za pyhon3 library"
