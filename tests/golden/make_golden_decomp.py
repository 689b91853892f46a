"""Golden vectors for the Tucker-2 decomposition, generated with the UNMODIFIED reference module
scripts/tensor_decomposition/decomposition.py (tensorly stubbed by oracle/decomp_oracle.py, whose HOOI restatement is
itself pinned by the reference's 6,329,941-parameter golden count). Run in the build container:
    python tests/golden/make_golden_decomp.py   ->  tests/golden/decomp_golden.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch import nn  # noqa: E402

from oracle import decomp_oracle  # noqa: E402

CASES = [  # (cout, cin, k, stride, true ranks, noise)
    (32, 16, 3, 1, (6, 5), 0.02),
    (64, 64, 3, 1, (12, 20), 0.03),
    (128, 64, 3, 2, (30, 16), 0.02),
    (96, 48, 3, 1, (48, 24), 0.05),
]


def make_layer(case, seed):
    cout, cin, k, s, (r0, r1), noise = case
    g = torch.Generator().manual_seed(seed)
    core = torch.randn(r0, r1, k, k, generator=g)
    a = torch.randn(cout, r0, generator=g) / r0 ** 0.5
    b = torch.randn(cin, r1, generator=g) / r1 ** 0.5
    w = torch.einsum("rskl,or,is->oikl", core, a, b) + noise * torch.randn(cout, cin, k, k, generator=g)
    conv = nn.Conv2d(cin, cout, k, s, k // 2, bias=True)
    conv.weight.data = w
    conv.bias.data = torch.randn(cout, generator=g)
    x = torch.randn(2, cin, 12, 12, generator=g)
    return conv, x


def main():
    ref = decomp_oracle.load_reference()
    out = []
    for i, case in enumerate(CASES):
        conv, x = make_layer(case, 100 + i)
        ranks = ref.estimate_ranks(conv)
        chain = ref.tucker_decomposition_conv_layer(conv)
        with torch.no_grad():
            y, y0 = chain(x), conv(x)
        out.append({"case": list(case[:4]) + [list(case[4]), case[5]], "seed": 100 + i, "ranks": [int(r) for r in ranks],
                    "chain_params": sum(p.numel() for p in chain.parameters()),
                    "mean_abs_diff": float((y - y0).abs().mean()), "out_abs_mean": float(y.abs().mean()),
                    "out_samples": [float(v) for v in y.flatten()[:: max(1, y.numel() // 16)][:16]]})
        print(out[-1]["case"], out[-1]["ranks"], out[-1]["mean_abs_diff"])
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "decomp_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
