"""Extracts the 91 MiMC7 round constants (public parameters: a keccak256 chain seeded with "mimc") from the reference's
circomlib copy into tests/golden/mimc7_constants.json.  Run in the build container, where /root/reference is mounted:
    python tests/golden/make_mimc7_constants.py
Source: /root/reference/interop/circuits/circomlib/circuits/mimc.circom:25-117 (`var c = [ ... ]` of MiMC7)."""
import json
import os
import re

SRC = "/root/reference/interop/circuits/circomlib/circuits/mimc.circom"
text = open(SRC).read()
body = text[text.index("var c = ["):]
body = body[:body.index("];")]
consts = [int(x) for x in re.findall(r"\b\d+\b", body)]
assert len(consts) == 91 and consts[0] == 0, len(consts)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mimc7_constants.json")
json.dump({"source": "interop/circuits/circomlib/circuits/mimc.circom:25-117 (MiMC7, 91 rounds)", "c": [str(c) for c in consts]}, open(out, "w"), indent=0)
print("wrote", out, len(consts))
