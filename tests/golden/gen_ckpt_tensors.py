"""Generates tests/golden/ckpt_tensors_V2_5.json from the reference's shipped GraphDef (run in the build container only):
the names the graph's Saver writes (save/SaveV2/tensor_names) and the shape of each of those variables (VarHandleOp `shape` attr).
    python tests/golden/gen_ckpt_tensors.py
"""
import json
import os
import re

SRC = "/root/reference/python/model/model_txt_V2_5.pb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ckpt_tensors_V2_5.json")

txt = open(SRC).read()
i = txt.index('name: "save/SaveV2/tensor_names"')
block = txt[i:txt.index("\nnode {", i)]
names = re.findall(r'string_val: "([^"]+)"', block)
shapes = {}
for m in re.finditer(r'node \{\n  name: "([^"]+)"\n  op: "VarHandleOp"(.*?)\n\}\nnode', txt, re.S):
    name, body = m.group(1), m.group(2)
    sm = re.search(r'key: "shape"\s+value \{\s+shape \{(.*?)\n      \}', body, re.S)
    shapes[name] = [int(d) for d in re.findall(r"size: (\d+)", sm.group(1))] if sm else []
out = {"source": "python/model/model_txt_V2_5.pb: save/SaveV2/tensor_names + VarHandleOp shapes", "tensors": [[n, shapes[n]] for n in names]}
json.dump(out, open(OUT, "w"), indent=0)
print(len(names), "tensors ->", OUT)
