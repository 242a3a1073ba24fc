"""CPU-only checks of the host side: the C-ABI library loads and exports exactly what the header
declares (no compute calls without a GPU), the public dataclasses mirror the reference, the segment
tables and the prefix/tail split, the packed winner key, and the product's refusal to run on CPU."""
import ctypes
import dataclasses
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "bimodal_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bma_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bimodalattack_b200 import native
    lib = native.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bimodal_b200.h but not exported"
    assert sorted(native.SIGNATURES) == names          # the ctypes table covers the whole header
    assert lib.bma_version() >= 100


def test_config_and_result_mirror_reference_fields():
    from bimodalattack_b200 import BimodalAttackConfig, BimodalAttackResult
    want = ["num_steps", "optim_str_init", "search_width", "batch_size", "topk", "n_replace", "buffer_size",
            "use_mellowmax", "mellowmax_alpha", "early_stop", "allow_non_ascii", "filter_ids",
            "add_space_before_target", "seed", "verbosity", "dynamic_search", "min_search_width", "alpha",
            "eps", "pgd_attack", "gcg_attack", "debug_output", "joint_eval", "experiment_folder",
            "images_folder", "pgd_after_gcg", "model"]                       # bimodal_attack.py:42-70
    assert [f.name for f in dataclasses.fields(BimodalAttackConfig)] == want
    c = BimodalAttackConfig()
    assert (c.num_steps, c.search_width, c.topk, c.n_replace, c.alpha, c.eps) == (250, 512, 256, 1, 0.01, 0.1)
    assert c.optim_str_init == " ".join(["x"] * 19) and c.filter_ids and not c.joint_eval
    assert [f.name for f in dataclasses.fields(BimodalAttackResult)] == [
        "best_loss", "best_string", "losses", "strings", "adversarial_suffixes", "model_outputs",
        "gradient_times", "sampling_times", "loss_times", "pgd_times", "total_times"]


def test_segment_orders_match_oracle_tables():
    from bimodalattack_b200.segments import segment_order
    from oracle import hotpath
    for mt in ("llava", "gemma3"):
        for mode, flags in [("pgd", dict(single=True)), ("gcg", dict(single=True)),
                            ("gcg", dict(no_joint_eval=True)), ("gcg", dict(no_target=True)),
                            ("gcg_pgd", dict(single=True)), ("gcg_pgd", dict(no_target=True)), ("gcg_pgd", {})]:
            assert segment_order(mode, mt, **flags) == hotpath.segment_order(mode, mt, **flags)
    with pytest.raises(AssertionError):
        segment_order("pgd", "llava", single=False)
    with pytest.raises(ValueError):
        segment_order("gcg", "llava")
    with pytest.raises(ValueError):
        segment_order("nope", "llava")


def test_plan_tail_splits_prefix_and_tail():
    from bimodalattack_b200.segments import plan_tail, segment_order
    d = 8
    parts = {n: torch.full((1, k, d), float(i)) for i, (n, k) in enumerate(
        [("before_img", 2), ("image", 5), ("before_suffix", 3), ("after", 4), ("target", 6)])}
    plan = plan_tail(segment_order("pgd", "llava", single=True), parts, n_optim=7)
    assert plan.n_prefix == 10 and plan.n_tail == 7 + 4 + 6
    assert plan.slot.tolist() == list(range(7)) + [-1] * 10
    assert torch.equal(plan.fixed[7:11], parts["after"][0]) and torch.equal(plan.fixed[11:], parts["target"][0])
    # Gemma order: the suffix comes before the image, so only before_img is shared
    plan = plan_tail(segment_order("pgd", "gemma3", single=True), parts, n_optim=7)
    assert plan.n_prefix == 2 and plan.n_tail == 7 + 3 + 5 + 4 + 6
    # GCG-only text order that starts with ... before | optim
    parts["before"] = torch.zeros((1, 3, d))
    plan = plan_tail(segment_order("gcg", "llava", no_joint_eval=True), parts, n_optim=7)
    assert plan.n_prefix == 3 and plan.slot.tolist()[:7] == list(range(7))


def test_packed_key_orders_like_argmin():
    from bimodalattack_b200.dist import pack_key
    from bimodalattack_b200.engine import unpack_key
    vals = [3.25, 3.25, -1.5, 0.0, 7.0, float("inf")]
    keys = [pack_key(v, i) for i, v in enumerate(vals)]
    best = min(keys)
    loss, idx = unpack_key(best)
    assert (loss, idx) == (-1.5, 2)
    assert all(0 <= k < 2 ** 63 for k in keys)                      # fits a signed int64 for the all-reduce
    assert unpack_key(min(keys[:2])) == (3.25, 0)                   # ties -> lowest index, like torch.argmin
    assert pack_key(float("nan"), 0) > pack_key(float("inf"), 5)


def test_product_refuses_cpu_tensors():
    from bimodalattack_b200 import native
    x = torch.zeros(4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        native.pgd_step_(x, x, x, 0.1, 0.1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        native.token_grad(x.view(1, 4).bfloat16(), x.view(1, 4).bfloat16())


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "bimodalattack_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn


def test_attack_buffer_semantics():
    from bimodalattack_b200.utils import AttackBuffer
    b = AttackBuffer(0)
    b.add(3.0, "a"); b.add(5.0, "b")
    assert b.get_best_ids() == "b" and b.get_lowest_loss() == 5.0          # size 0 keeps the latest (:97-99)
    b = AttackBuffer(2)
    b.add(3.0, "a"); b.add(1.0, "b"); b.add(2.0, "c")
    assert [e[1] for e in b.buffer] == ["b", "c"] and b.get_highest_loss() == 2.0


def test_png_writer_matches_reference_conversion(tmp_path):
    """_save_image (reference :1312-1317) runs off-thread here; the bytes must be the reference's
    ``(arr * 255).astype(uint8)`` in HWC order."""
    import numpy as np
    import torch
    from PIL import Image
    from bimodalattack_b200.attack import BimodalAttack

    class Host:
        _png = None
    h = Host()
    imgs = [torch.rand((1, 3, 9, 14), generator=torch.Generator().manual_seed(i)) for i in range(4)]
    for i, img in enumerate(imgs):
        BimodalAttack._save_image(h, img, str(tmp_path / f"{i}.png"))
    BimodalAttack.flush_images(h)
    for i, img in enumerate(imgs):
        want = (img.squeeze(0).numpy().transpose(1, 2, 0) * 255).astype(np.uint8)
        assert (np.array(Image.open(tmp_path / f"{i}.png")) == want).all()


def test_patch_embed_gemm_equals_conv():
    """The unfold+GEMM forward installed on the tower's patch-embedding convolution (and its input
    gradient) equals the convolution; restore_patch_embed undoes it."""
    import torch
    from bimodalattack_b200.attack import install_patch_embed_gemm, restore_patch_embed
    torch.manual_seed(0)
    for bias in (False, True):
        conv = torch.nn.Conv2d(3, 24, kernel_size=14, stride=14, bias=bias)
        net = torch.nn.Sequential(conv)
        x = torch.rand(2, 3, 56 + 5, 42, requires_grad=True)           # ragged height: the conv drops the remainder
        want = net(x)
        g_want, = torch.autograd.grad(want.square().sum(), x)
        assert install_patch_embed_gemm(net) == 1 and install_patch_embed_gemm(net) == 0
        got = net(x)
        g_got, = torch.autograd.grad(got.square().sum(), x)
        assert got.shape == want.shape and torch.allclose(got, want, atol=1e-5)
        assert torch.allclose(g_got, g_want, atol=1e-4)
        import copy
        twin = copy.deepcopy(net).double()                               # a copy runs on ITS OWN weights
        assert twin(x.double()).dtype == torch.float64
        restore_patch_embed(net)
        assert torch.equal(net(x), want)
    other = torch.nn.Sequential(torch.nn.Conv2d(3, 8, kernel_size=3, stride=1, padding=1))
    assert install_patch_embed_gemm(other) == 0                          # overlapping windows: left alone


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle on the host cores) prints ONE JSON line with the keys
    the driver reads; here on the CPU-sized model of BASELINE.json config 1."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--model", "tiny",
                        "--search-width", "64", "--cpu-sample", "4", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gcg_pgd_candidates_scored_per_sec"
    assert d["unit"] == "candidates/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["search_width"] == 64
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "candidate" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
