"""CPU: the oracle (oracle/hotpath.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/*.json, generated in the build container by `python -m oracle.make_golden`).
Same torch ops in the same order on the same seeds, so suffix ids must match exactly; losses are
compared to 1e-5 to tolerate BLAS differences between hosts."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases, hotpath

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


RUNS = _load("reference_runs.json")["runs"]


@pytest.mark.parametrize("run", RUNS, ids=[f"{r['family']}-{r['mode']}" for r in RUNS])
def test_oracle_replays_reference_run(run):
    from bimodalattack_b200.config import BimodalAttackConfig
    model, tok, proc, norm, image = cases.build(run["family"])
    kw = cases.config_kwargs(run["mode"])
    cfg = BimodalAttackConfig(**kw)
    atk = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
    log = atk.run([{"role": "user", "content": cases.PROMPT}], cases.TARGET,
                  image.clone() if kw["pgd_attack"] else None)
    strings = [tok.batch_decode(torch.tensor([ids]))[0] for ids in log.suffix_ids]
    assert strings == run["strings"]
    assert np.allclose(log.losses, run["losses"], rtol=0, atol=1e-5)
    if run["image_u8_checks"] is not None:
        arr = (atk.final_image.squeeze(0).detach().numpy().transpose(1, 2, 0) * 255).astype(np.uint8)
        got = arr.astype(np.int64).sum(axis=(0, 1)).tolist() + [int(arr[3, 5, 1]), int(arr[20, 11, 2])]
        assert got == run["image_u8_checks"]


def test_oracle_sampler_matches_reference_vector():
    v = _load("reference_functions.json")["sample"]
    g = torch.Generator().manual_seed(v["seed_inputs"])
    grad = torch.randn((v["L"], v["V"]), generator=g)
    ids = torch.randint(0, v["V"], (v["L"],), generator=g)
    torch.manual_seed(v["seed_rng"])
    got = hotpath.sample_ids_from_grad(ids, grad, v["sw"], v["topk"], v["n_replace"], torch.tensor(v["banned"]))
    assert got.tolist() == v["new_ids"]


def test_oracle_pgd_matches_reference_vector():
    v = _load("reference_functions.json")["pgd"]
    g = torch.Generator().manual_seed(v["seed"])
    shape = tuple(v["shape"])
    img = torch.rand(shape, generator=g)
    orig = (img + 0.3 * (torch.rand(shape, generator=g) - 0.5)).clamp(0, 1)
    gr = torch.randn(shape, generator=g)
    gr.view(-1)[::5] = 0.0
    out = hotpath.pgd_step(img, v["eps"], v["alpha"], gr, orig)
    assert [float(x).hex() for x in out.detach().view(-1).tolist()] == v["out_hex"]


def test_filter_ids_matches_reference_vector():
    """filter_ids (:166-186) with a tokenizer whose round trip really fails: the oracle's per-row loop and the
    product's single batched tokenizer call keep exactly the rows the reference keeps."""
    from bimodalattack_b200.attack import filter_ids as product_filter_ids
    v = _load("reference_functions.json")["filter"]
    ids, tok = cases.filter_case()
    want = torch.tensor(v["kept"])
    assert 0 < want.shape[0] < v["rows"]
    assert torch.equal(hotpath.filter_ids(ids.clone(), tok), want)
    assert torch.equal(product_filter_ids(ids.clone(), tok), want)


def test_nonascii_tokens_match_reference_vector():
    """get_nonascii_toks (reference utils.py:14-33): the not-allowed mask of allow_non_ascii=False."""
    from bimodalattack_b200 import synthetic as syn
    from bimodalattack_b200.utils import get_nonascii_toks
    v = _load("reference_functions.json")["nonascii"]
    assert get_nonascii_toks(syn.make_tokenizer(v["vocab"])).tolist() == v["ids"] and len(v["ids"]) > 0
