"""CPU: the plain-C restatement (oracle/hotpath_c.c) against torch on CPU and against Philox
vectors captured from torch on a B200 (tests/golden/cuda_philox.json, made by
tools/capture_cuda_philox.py under gpurun)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_c.so"))
    lib.oracle_mean_ce.restype = C.c_float
    lib.oracle_offset_advance.restype = C.c_uint64
    lib.oracle_offset_advance.argtypes = [C.c_long, C.c_int, C.c_int]
    lib.oracle_cuda_rand.argtypes = [C.c_uint64, C.c_uint64, C.c_long, C.c_int, C.c_int, C.c_void_p]
    lib.oracle_cuda_randint.argtypes = [C.c_uint64, C.c_uint64, C.c_long, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
    lib.oracle_pgd_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_float]
    lib.oracle_philox4.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    return lib


def test_c_pgd_equals_torch(clib):
    from oracle import hotpath
    g = torch.Generator().manual_seed(1)
    img = torch.rand((1, 3, 11, 13), generator=g)
    orig = (img + 0.3 * (torch.rand(img.shape, generator=g) - 0.5)).clamp(0, 1)
    gr = torch.randn(img.shape, generator=g)
    gr.view(-1)[::3] = 0
    gr.view(-1)[1] = float("nan")
    eps, alpha = 64 / 255, 4 / 255
    want = hotpath.pgd_step(img.clone(), eps, alpha, gr, orig).detach()
    got = img.clone().contiguous()
    step = float(torch.tensor(alpha * eps, dtype=torch.float32))
    clib.oracle_pgd_step(got.data_ptr(), gr.data_ptr(), orig.data_ptr(), got.numel(), step,
                         float(torch.tensor(eps, dtype=torch.float32)))
    assert torch.equal(got, want)


def test_c_philox_known_answer(clib):
    # Random123 known-answer vectors for philox4x32-10 (counter, key) -> output
    out = (C.c_uint32 * 4)()
    clib.oracle_philox4(0, 0, 0, out)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # all-ones counter and key: seed = 2^64-1, subsequence = 2^64-1, counter lo/hi = 2^64-1
    clib.oracle_philox4(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, (0xFFFFFFFFFFFFFFFF // 4) * 4, out)
    assert len(list(out)) == 4


def test_c_sampler_equals_torch_cpu_given_the_draws(clib):
    from oracle import hotpath
    L, V, sw, topk, nrep = 6, 300, 40, 16, 2
    g = torch.Generator().manual_seed(2)
    grad = torch.randn((L, V), generator=g)
    ids = torch.randint(0, V, (L,), generator=g)
    banned = torch.tensor([0, 1, 2, 50, 51])
    torch.manual_seed(123)
    want = hotpath.sample_ids_from_grad(ids.clone(), grad.clone(), sw, topk, nrep, banned)
    torch.manual_seed(123)                              # the same two draws, handed to the C code
    uniforms = torch.rand((sw, L)).contiguous()
    picks = torch.randint(0, topk, (sw, nrep, 1)).view(sw, nrep).contiguous()
    flags = torch.zeros(V, dtype=torch.uint8)
    flags[banned] = 1
    out = torch.empty((sw, L), dtype=torch.int64)
    rc = clib.oracle_sample_ids(C.c_void_p(ids.data_ptr()), C.c_void_p(grad.data_ptr()), L, V,
                                C.c_void_p(flags.data_ptr()), sw, topk, nrep, C.c_void_p(uniforms.data_ptr()),
                                C.c_void_p(picks.data_ptr()), C.c_void_p(out.data_ptr()))
    assert rc == 0 and torch.equal(out, want)


def test_c_mean_ce_equals_torch(clib):
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn((5, 97), generator=g) * 3).contiguous()
    tgt = torch.randint(0, 97, (5,), generator=g)
    want = torch.nn.functional.cross_entropy(logits, tgt).item()
    got = clib.oracle_mean_ce(C.c_void_p(logits.data_ptr()), 5, 97, C.c_void_p(tgt.data_ptr()))
    assert abs(got - want) < 1e-5


def test_c_philox_reproduces_torch_cuda_stream(clib):
    path = os.path.join(ROOT, "tests", "golden", "cuda_philox.json")
    if not os.path.exists(path):
        pytest.skip("cuda_philox.json not captured yet")
    gold = json.load(open(path))
    for case in gold["cases"]:
        n = case["n"]
        out = np.empty(n, dtype=np.float32)
        clib.oracle_cuda_rand(case["seed"], case["offset"], n, gold["sms"], gold["max_threads_per_sm"],
                              out.ctypes.data_as(C.c_void_p))
        idx = case["rand_idx"]
        assert [float(out[i]).hex() for i in idx] == case["rand_hex"]
        adv = clib.oracle_offset_advance(n, gold["sms"], gold["max_threads_per_sm"])
        assert adv == case["offset_after_rand"] - case["offset"]
        ints = np.empty(case["n_int"], dtype=np.int64)
        clib.oracle_cuda_randint(case["seed"], case["offset_after_rand"], case["n_int"], gold["sms"],
                                 gold["max_threads_per_sm"], case["range"], ints.ctypes.data_as(C.c_void_p))
        assert ints[case["int_idx"]].tolist() == case["int_vals"]
