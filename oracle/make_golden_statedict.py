"""TEST INFRASTRUCTURE -- freeze the key / shape / dtype list of the REFERENCE model's ``state_dict()`` (what
``save_model`` writes into ``epoch-N.pt`` as ``model_state_dict``, Code_Cached/data_utils/utils.py:104-110, and what
run.py:234-243 loads back) -> tests/golden/state_dict_keys.json.  One subprocess per source tree (both own the module name
``model``).  Run in the build container only:

    python -m oracle.make_golden_statedict
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")

CASES = {
    "Code_Cached": {},
    "Code_Cached_Asym": dict(asym=True, d_text=1024, d_img=768, layers_text=25, layers_img=13,
                             bert_list="1,3,5,7,9,11,13,15,17,19,21,23", vit_list="1,3,5,7,9,11"),
}


def run_case(tree):
    sys.dont_write_bytecode = True
    sys.path.insert(0, ROOT)
    from oracle.synthetic import PathConfig, make_args
    cfg = PathConfig(**CASES[tree])
    sys.path.insert(0, os.path.join("/root/reference", tree))
    from torch import nn
    from model.model import ModelMM, IISANAdaptedMModel          # the reference's own code
    args = make_args(cfg)

    class ImgStub(nn.Module):
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(cfg.d_img, cfg.embedding_dim)

    m = ModelMM(args, 50, True, ImgStub(), nn.Identity(), [1.0] * 51)
    m.mm_encoder = IISANAdaptedMModel(m.mm_encoder, args)        # Code_Cached/run.py:182-183
    sd = m.state_dict()
    print(json.dumps({"cfg": cfg.to_dict(), "entries": [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]}))


def main():
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
        return
    out = {}
    for tree in CASES:
        r = subprocess.run([sys.executable, "-m", "oracle.make_golden_statedict", tree], cwd=ROOT, check=True, capture_output=True, text=True)
        out[tree] = json.loads(r.stdout.strip().splitlines()[-1])
        print(tree, len(out[tree]["entries"]), "state_dict entries")
    with open(OUT, "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
