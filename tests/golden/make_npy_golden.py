"""Extracts golden vectors for the NPY wire format from the reference's own test data (src/io/npy_test_data.js, generated
there by NumPy: base64 of a .npy file + the expected NDArray literal) into tests/golden/npy_golden.json.

Run in the build container only (reads /root/reference):  python tests/golden/make_npy_golden.py
Keeps every int32 / float32 / float64 example of at most 24 elements (both byte orders, C and Fortran order)."""
import json
import os
import re

SRC = "/root/reference/src/io/npy_test_data.js"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "npy_golden.json")

text = open(SRC).read()
pat = re.compile(r'yield \["([A-Za-z0-9+/=]+)",\s*new NDArray\( Int32Array\.of\(([^)]*)\), ARRAY_TYPES\["(\w+)"\]\.of\(([^)]*)\) \)\]')
items = []
for b64, shape, dtype, vals in pat.findall(text):
    if dtype not in ("int32", "float32", "float64"):
        continue
    shape = [int(s) for s in shape.split(",") if s.strip()]
    vals = [v.strip() for v in vals.split(",") if v.strip()]
    if len(vals) > 24:
        continue
    items.append({"b64": b64, "dtype": dtype, "shape": shape, "data": vals})  # values kept as the JS literals
items = items[::3]  # one in three is plenty (both byte orders, both memory orders and all three dtypes stay covered)
json.dump({"source": "nd4js src/io/npy_test_data.js", "items": items}, open(OUT, "w"), indent=0)
print(len(items), "examples ->", OUT, os.path.getsize(OUT), "bytes")
