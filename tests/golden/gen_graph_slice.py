"""Extracts the INFERENCE slice of the reference's shipped GraphDef (python/model/model_txt_V2_5.pb, the text form of the graph
neural_network/alphazero_nn.cpp loads) into tests/golden/graph_V2_5_inference.json: every node that output_policy / output_value
depend on with input_training = false, the If ops replaced by the body of their else_branch function (inlined, names prefixed).
The JSON is what oracle/graphdef_oracle.py executes; this script is the only thing that reads /root/reference.

A second file, graph_V2_5_training.json, is the same for a training call (input_training = true, then_branch of every If): the nodes
the two losses, the minimised total and the moving-average updates depend on, plus the optimizer's wiring and constants.

    python tests/golden/gen_graph_slice.py [/root/reference/python/model/model_txt_V2_5.pb] [inference.json] [training.json]
"""
import json
import os
import re
import sys

TOKEN = re.compile(r'\s*(?:(#[^\n]*)|([A-Za-z_][A-Za-z0-9_.]*)|("(?:[^"\\]|\\.)*")|(-?[0-9.][0-9.eE+\-]*|-?inf|nan)|([{}:<>\[\],]))')


def tokenize(text):
    pos, n = 0, len(text)
    while True:
        m = TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip():
                raise ValueError("cannot tokenize at %d: %r" % (pos, text[pos:pos + 40]))
            return
        pos = m.end()
        if m.group(1):
            continue
        yield m


def unescape(s):
    """C-escaped text-proto string -> list of byte values"""
    out, i, s = [], 0, s[1:-1]
    while i < len(s):
        c = s[i]
        if c != "\\":
            out.extend(c.encode("utf-8")); i += 1; continue
        c = s[i + 1]
        if c in "01234567":
            j = i + 1
            while j < len(s) and j < i + 4 and s[j] in "01234567":
                j += 1
            out.append(int(s[i + 1:j], 8)); i = j
        elif c == "x":
            out.append(int(s[i + 2:i + 4], 16)); i += 4
        else:
            out.append({"n": 10, "r": 13, "t": 9, "\\": 92, '"': 34, "'": 39, "a": 7, "b": 8, "f": 12, "v": 11}[c]); i += 2
    return out


def parse(text):
    """text proto -> nested {field: [values]}; message values are dicts, strings are byte lists, numbers stay text"""
    root, stack, field = {}, [], None
    cur = root
    for m in tokenize(text):
        ident, string, number, punct = m.group(2), m.group(3), m.group(4), m.group(5)
        if punct == "{":
            child = {}
            cur.setdefault(field, []).append(child)
            stack.append(cur); cur = child; field = None
        elif punct == "}":
            cur = stack.pop(); field = None
        elif punct == ":":
            continue
        elif field is None:
            field = ident
        else:
            cur.setdefault(field, []).append(("s", unescape(string)) if string else ("n", number) if number else ("e", ident))
            field = None
    return root


def text_of(v):
    return bytes(v[1]).decode("utf-8") if v[0] == "s" else v[1]


def attr_value(av):
    """AttrValue message -> JSON-able python value"""
    if "s" in av: return {"s": text_of(av["s"][0])}
    if "i" in av: return {"i": int(text_of(av["i"][0]))}
    if "f" in av: return {"f": float(text_of(av["f"][0]))}
    if "b" in av: return {"b": text_of(av["b"][0]) == "true"}
    if "type" in av: return {"type": text_of(av["type"][0])}
    if "func" in av: return {"func": text_of(av["func"][0]["name"][0])}
    if "shape" in av: return {"shape": [int(text_of(d["size"][0])) if "size" in d else 0 for d in av["shape"][0].get("dim", [])]}
    if "tensor" in av:
        t = av["tensor"][0]
        out = {"dtype": text_of(t["dtype"][0]), "shape": [int(text_of(d["size"][0])) if "size" in d else 0 for d in t.get("tensor_shape", [{}])[0].get("dim", [])]}
        if "tensor_content" in t: out["content"] = t["tensor_content"][0][1]
        for k in ("float_val", "int_val", "bool_val", "int64_val"):
            if k in t: out[k] = [text_of(x) for x in t[k]]
        return {"tensor": out}
    if "list" in av:
        l = av["list"][0]
        for k in ("i", "s", "type", "f", "b"):
            if k in l: return {"list_" + k: [text_of(x) for x in l[k]]}
        return {"list": []}
    return {}


def node_json(nd, prefix=""):
    return {"name": prefix + text_of(nd["name"][0]), "op": text_of(nd["op"][0]), "input": [text_of(x) for x in nd.get("input", [])],
            "attr": {text_of(a["key"][0]): attr_value(a["value"][0]) for a in nd.get("attr", [])}}


def main(src, dst, outputs=("output_policy", "output_value"), branch="else_branch", n_if_outputs=1):
    """branch: which function of every If / StatelessIf runs (else_branch: input_training = false, then_branch: true);
    n_if_outputs: how many leading outputs of an If are kept (1: the normalised tensor; 3: + batch mean and variance)"""
    g = parse(open(src).read())
    nodes = {text_of(n["name"][0]): n for n in g["node"]}
    funcs = {text_of(f["signature"][0]["name"][0]): f for f in g["library"][0]["function"]}
    out, done = [], set()

    def tensor_node(t):
        t = t.lstrip("^")
        return t.split(":")[0]

    def visit(name):
        if name in done:
            return
        done.add(name)
        nd = nodes[name]
        op = text_of(nd["op"][0])
        j = node_json(nd)
        if op in ("If", "StatelessIf"):
            # input_training is fed false: only the else branch runs.  Inline it: function argument k = If input k + 1;
            # the If node itself becomes an "IfOutputs" node whose inputs are the function's return tensors (in output_arg order)
            fn = funcs[j["attr"][branch]["func"]]
            sig = fn["signature"][0]
            args = [text_of(a["name"][0]) for a in sig.get("input_arg", [])]
            prefix = name + "/" + branch.split("_")[0] + "/"
            argmap = {a: j["input"][k + 1] for k, a in enumerate(args)}
            body = {text_of(n["name"][0]): n for n in fn.get("node_def", [])}
            ret = {text_of(r["key"][0]): text_of(r["value"][0]) for r in fn.get("ret", [])}

            def remap(t):                       # function-local tensor name "node:out_name:idx" / argument name -> graph-level name
                base = t.split(":")[0]
                if base in argmap:
                    return argmap[base]
                parts = t.split(":")
                return prefix + parts[0] + ":" + (parts[2] if len(parts) == 3 else "0") + ("#" + parts[1] if len(parts) == 3 else "")

            inner_done = set()

            def visit_inner(local):
                if local in inner_done or local in argmap:
                    return
                inner_done.add(local)
                jn = node_json(body[local], prefix)
                for t in jn["input"]:
                    visit_inner(t.split(":")[0])
                jn["input"] = [remap(t) for t in jn["input"]]
                for t in jn["input"]:
                    if not t.startswith(prefix):
                        visit(tensor_node(t))
                out.append(jn)
            # only the leading outputs are consumed on the paths of interest (0: the tensor; 1, 2: batch mean / variance)
            outs = [text_of(o["name"][0]) for o in sig["output_arg"]][:n_if_outputs]
            for o in outs:
                visit_inner(ret[o].split(":")[0])
            visit(tensor_node(j["input"][0]))            # the predicate, kept so that the slice shows what selects the branch
            out.append({"name": name, "op": "IfElseOutput" if branch == "else_branch" else "IfThenOutput",
                        "input": [remap(ret[o]) if ret[o].split(":")[0] not in argmap else argmap[ret[o].split(":")[0]] for o in outs] + [j["input"][0]],
                        "attr": {"else_branch": j["attr"]["else_branch"], "then_branch": j["attr"]["then_branch"]}})
            return
        for t in j["input"]:
            if not t.startswith("^"):
                visit(tensor_node(t))
        j["input"] = [t for t in j["input"] if not t.startswith("^")]
        out.append(j)

    for o in outputs:
        visit(o)
    json.dump({"source": "python/model/" + os.path.basename(src), "outputs": list(outputs), "branch": branch, "nodes": out}, open(dst, "w"), indent=0)
    ops = {}
    for n in out:
        ops[n["op"]] = ops.get(n["op"], 0) + 1
    print("wrote %s: %d nodes" % (dst, len(out)), ops)


def training_slice(src, dst):
    """input_training = true: the two losses, the total loss the optimizer minimises, the value every moving statistic is decreased by
    (AssignMovingAvg: AssignSubVariableOp(variable, <node>)), and what the 45 ResourceApplyAdam ops are wired to"""
    g = parse(open(src).read())
    nodes = {text_of(n["name"][0]): n for n in g["node"]}
    updates, adam_vars, adam_inputs, nesterov = [], [], set(), set()
    for name, nd in nodes.items():
        op = text_of(nd["op"][0])
        ins = [text_of(x) for x in nd.get("input", [])]
        if op == "AssignSubVariableOp" and "AssignMovingAvg" in name:
            updates.append([ins[0], ins[1]])
        if op == "ResourceApplyAdam":
            adam_vars.append(ins[0])
            adam_inputs.add(tuple(ins[5:9]))          # lr, beta1, beta2, epsilon (inputs 3 and 4 read the beta1_power / beta2_power variables)
            assert [text_of(nodes[ins[k]]["input"][0]) for k in (3, 4)] == ["beta1_power", "beta2_power"] and ins[1:3] == [ins[0] + "/optimize", ins[0] + "/optimize_1"]
            attrs = {text_of(a["key"][0]): attr_value(a["value"][0]) for a in nd.get("attr", [])}
            nesterov.add(bool(attrs.get("use_nesterov", {"b": False})["b"]))
            # the gradient fed to the update is d(total loss)/d(variable): its producer lives under gradients/ of the total loss
    assert len(adam_inputs) == 1 and nesterov == {False}
    hyper = list(adam_inputs)[0]
    outputs = ["softmax_cross_entropy_loss/value", "mean_squared_error/value", "add_6"] + [u[1] for u in sorted(updates)]
    main(src, dst, outputs=outputs, branch="then_branch", n_if_outputs=3)
    d = json.load(open(dst))

    def const_value(name):
        t = attr_value([a for a in nodes[name]["attr"] if text_of(a["key"][0]) == "value"][0]["value"][0])["tensor"]
        return float(t["float_val"][0])
    grad_of_total = [n for n in nodes if n.startswith("gradients/add_6_grad/")]           # tf.gradients(add_6, ...): add_6 is what is minimised
    d["adam"] = {"variables": sorted(adam_vars), "learning_rate": const_value(hyper[0]), "beta1": const_value(hyper[1]), "beta2": const_value(hyper[2]),
                 "epsilon": const_value(hyper[3]), "use_nesterov": False, "hyper_nodes": list(hyper), "slots": ["<variable>/optimize", "<variable>/optimize_1"],
                 "minimised": "add_6" if grad_of_total else None}
    d["moving_average_updates"] = sorted(updates)
    json.dump(d, open(dst, "w"), indent=0)
    print("training slice: %d moving-average updates, %d Adam variables, lr %g beta1 %g beta2 %g eps %g"
          % (len(updates), len(adam_vars), d["adam"]["learning_rate"], d["adam"]["beta1"], d["adam"]["beta2"], d["adam"]["epsilon"]))


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/python/model/model_txt_V2_5.pb"
    main(src, sys.argv[2] if len(sys.argv) > 2 else os.path.join(here, "graph_V2_5_inference.json"))
    training_slice(src, sys.argv[3] if len(sys.argv) > 3 else os.path.join(here, "graph_V2_5_training.json"))
