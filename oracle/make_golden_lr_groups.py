"""TEST INFRASTRUCTURE -- freeze the learning-rate group of every parameter as the REFERENCE's training script routes it
(Code_Cached/run.py:260-307: the ``if use_modal:`` block of ``train`` that fills image_net_params / text_encoder_params /
recsys_params / adapter_cv_params / adapter_text_params by substring tests on the parameter name and builds optim.Adam) ->
tests/golden/lr_groups.json.

The block sits inside ``train`` and cannot be imported on its own (run.py needs lmdb / loralib), so its source text is taken
from /root/reference with ``ast`` and executed as is over the parameter names of the reference's own model
(state_dict_keys.json lists them); ``optim.Adam`` is replaced by a recorder.  Nothing is copied into the repository.
Run in the build container only:

    python -m oracle.make_golden_lr_groups
"""
import ast
import json
import os
import types

REF = "/root/reference/Code_Cached/run.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "lr_groups.json")
LRS = dict(lr=2e-4, adapter_cv_lr=1e-4, adapter_bert_lr=1e-4, fine_tune_lr_image=1e-4, fine_tune_lr_text=5e-5)   # scripts/run_IISAN.py:30-43


def routing_block():
    tree = ast.parse(open(REF).read())
    train = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train")
    # the `if use_modal:` statement whose body starts with `image_net_params = []`
    for node in train.body:
        if isinstance(node, ast.If) and isinstance(node.body[0], ast.Assign) and \
                getattr(node.body[0].targets[0], "id", "") == "image_net_params":
            stmts = []
            for s in node.body:                      # up to and including the statement that builds optim.Adam (run.py:294-307);
                stmts.append(s)                      # what follows only logs the groups
                if "optim.Adam" in ast.unparse(s):
                    break
            return ast.Module(body=stmts, type_ignores=[])
    raise RuntimeError("routing block not found")


class P:                       # stands in for a Parameter: identity is all the block uses
    def __init__(self, name):
        self.name, self.requires_grad = name, True


def main():
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    out = {}
    for tree, adding in (("Code_Cached", "all"), ("Code_Cached_Asym", "all")):
        names = [e[0] for e in keys[tree]["entries"]]
        params = [P(n) for n in names]
        rec = {}

        class Adam:
            def __init__(self, groups, **kw):
                rec["groups"] = groups

        model = types.SimpleNamespace(module=types.SimpleNamespace(named_parameters=lambda: [(p.name, p) for p in params]))
        ns = {"model": model, "args": types.SimpleNamespace(adding_adapter_to=adding, **LRS), "optim": types.SimpleNamespace(Adam=Adam)}
        exec(compile(routing_block(), REF, "exec"), ns)
        lr_of = {}
        for g in rec["groups"]:
            for p in g["params"]:
                assert p.name not in lr_of
                lr_of[p.name] = g["lr"]
        assert set(lr_of) == set(names)
        out[tree] = {"lrs": LRS, "lr_of": lr_of, "group_sizes": [len(g["params"]) for g in rec["groups"]]}
        print(tree, out[tree]["group_sizes"])
    json.dump(out, open(OUT, "w"))


if __name__ == "__main__":
    main()
