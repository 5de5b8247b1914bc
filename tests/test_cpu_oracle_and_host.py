"""CPU tier (-m "not gpu"): oracle vs golden fixtures, host logic, ABI export check."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import cases
import styletts2_oracle as O
from util import GOLD, ROOT, apply_patch, golden, oracle_sds, ref_shapes


@pytest.mark.parametrize("cname", list(cases.E2E_CASES))
def test_oracle_reproduces_reference_e2e_fixture(cname):
    """The committed fixtures are outputs of the UNMODIFIED reference (oracle/make_golden.py); the
    oracle restatement must reproduce them from seeds alone: integer durations bit-exact, floats 1e-4."""
    case = cases.E2E_CASES[cname]
    g = golden(cname)
    mcfg = cases.MODEL_CFGS[case["model"]]
    sds = oracle_sds(case["model"])
    tokens, lengths, bert_dur, noise, ref_s = cases.e2e_inputs(case)
    rng = cases.ReplayRNG(case["seed"])
    B, L = case["B"], g["wav"].shape[-1]
    inj = dict(step_noises=[rng.step_noise(i, (B, 1, 256)) for i in range(case["steps"] - 1)],
               rand_ini=rng.rand_ini((B, 9)), sine_noise=rng.sine_noise((B, L, 9)))
    forced = torch.from_numpy(g["forced_dur"]).float()
    with torch.no_grad():
        if mcfg["decoder"]["type"] == "istftnet":
            har = O.istftnet_har(torch.from_numpy(g["F0"]), O.sub(sds["decoder"], "generator"), mcfg["decoder"],
                                 inj["rand_ini"], inj["sine_noise"])
            inj["har"] = apply_patch(har, g["har_patch_idx"], g["har_patch_val"])
        out = O.synthesize(sds, mcfg, tokens, lengths, bert_dur, noise, diffusion_steps=case["steps"],
                           embedding_scale=case["embedding_scale"], ref_s=ref_s, rng=inj, forced_durations=forced)
    assert np.array_equal(out["pred_dur"].numpy().astype(np.int32), g["pred_dur"]), "integer durations must be bit-exact"
    for k, tol in [("s_pred", 1e-5), ("logits", 1e-4), ("F0", 1e-3), ("N", 1e-4)]:
        d = float((out[k] - torch.from_numpy(g[k])).abs().max())
        assert d <= tol, (k, d)
    dw = float((out["wav"].squeeze(1) - torch.from_numpy(g["wav"])).abs().max())
    assert dw <= 1e-4, dw


@pytest.mark.parametrize("cname", list(cases.DECODER_CASES))
def test_oracle_reproduces_reference_decoder_fixture(cname):
    case = cases.DECODER_CASES[cname]
    g = golden(cname)
    mcfg = cases.MODEL_CFGS[case["model"]]
    sds = oracle_sds(case["model"], ("decoder",))
    asr, f0, n, s = cases.decoder_inputs(case)
    rng = cases.ReplayRNG(case["seed"])
    L = g["wav"].shape[-1]
    ri, sn = rng.rand_ini((case["B"], 9)), rng.sine_noise((case["B"], L, 9))
    har = None
    with torch.no_grad():
        if mcfg["decoder"]["type"] == "istftnet":
            har = apply_patch(O.istftnet_har(f0, O.sub(sds["decoder"], "generator"), mcfg["decoder"], ri, sn),
                              g["har_patch_idx"], g["har_patch_val"])
        wav = O.decoder(asr, f0, n, s, sds["decoder"], mcfg["decoder"], ri, sn, har).squeeze(1)
    assert float((wav - torch.from_numpy(g["wav"])).abs().max()) <= 1e-4


def test_state_dict_schema_matches_reference():
    """Drop-in boundary: our build_model exposes the reference's state-dict keys/shapes."""
    from styletts2_b200.models import build_model, recursive_munch
    for name in cases.MODEL_CFGS:
        m = build_model(recursive_munch(cases.MODEL_CFGS[name]))
        ref = ref_shapes(name)
        for k in ref:
            mine = {n: list(v.shape) for n, v in m[k].state_dict().items()}
            assert mine == ref[k], (name, k)


def test_build_model_container_keys():
    from styletts2_b200.models import build_model, recursive_munch
    m = build_model(recursive_munch(cases.MODEL_CFGS["ljspeech"]))
    assert list(m.keys()) == ["bert", "bert_encoder", "predictor", "decoder", "text_encoder", "predictor_encoder",
                              "style_encoder", "diffusion", "text_aligner", "pitch_extractor", "mpd", "msd", "wd"]
    _ = [m[k].eval() for k in m]  # notebook cell 10


def test_karras_schedule_and_adpm2_scalars_match_oracle():
    from styletts2_b200.diffusion import ADPM2Sampler, KarrasSchedule
    for K in (3, 5, 10, 50):
        s1 = KarrasSchedule(sigma_min=0.0001, sigma_max=3.0, rho=9.0)(K, "cpu")
        s2 = O.karras_sigmas(K)
        assert torch.equal(s1, s2)
    s = O.karras_sigmas(5)
    assert abs(float(s[0]) - 3.0) < 1e-6 and float(s[-1]) == 0.0 and abs(float(s[-2]) - 1e-4) < 1e-9
    up, down, mid = ADPM2Sampler().get_sigmas(s[0], s[1])
    assert 0 < down < float(s[1]) and float(s[1]) < float(mid) < float(s[0]) and up > 0


def test_keyed_weights_are_deterministic_and_alias_consistent():
    sds = oracle_sds("ljspeech", ("diffusion",))["diffusion"]
    for k, v in sds.items():
        if k.startswith("unet."):
            assert torch.equal(v, sds["diffusion.net." + k[len("unet."):]])
    from styletts2_b200.synthetic import keyed_tensor
    assert torch.equal(keyed_tensor("decoder.encode.conv1.weight_v", (4, 3, 3)), keyed_tensor("decoder.encode.conv1.weight_v", (4, 3, 3)))


def test_c_abi_library_exports_every_declared_symbol(lib_built):
    """include/styletts2_b200.h <-> libstyletts2_b200.so <-> ctypes table (no compute calls: no GPU here)."""
    hdr = open(os.path.join(ROOT, "include", "styletts2_b200.h")).read()
    declared = set(re.findall(r"\b(st2_[a-zA-Z0-9_]+)\s*\(", hdr))
    declared -= {"st2_conv_args", "st2_rows_args"}
    from styletts2_b200 import lib as L
    assert declared == set(L.SIGNATURES), (declared ^ set(L.SIGNATURES))
    so = ctypes.CDLL(lib_built)
    for name in declared:
        assert hasattr(so, name), name
    lib = L.load()
    assert lib.st2_abi_version() == L.ABI_VERSION == 2
    assert L.launch_count() == 0


def test_product_path_fails_loudly_without_cuda(lib_built):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from styletts2_b200 import ops
    with pytest.raises(RuntimeError):
        ops.scale(torch.zeros(4), 2.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "styletts2_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "styletts2_oracle" not in src and "import oracle" not in src and "ref_import" not in src, fn


def test_utterance_sharding_plan():
    from styletts2_b200.parallel import shard_range
    B = 64
    for W in (1, 2, 4, 8):
        spans = [shard_range(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        assert all(e - s == B // W for s, e in spans)
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_oracle_long_form_carry_over_properties():
    """LFinference restatement: t=0 ignores s_prev; carrying a sentence's own unblended style back in is a fixed point
    (single speaker); the multispeaker carry is cat([ref, s]) after the ref_s blend."""
    case = dict(model="ljspeech", B=1, N=6, seed=3)
    tokens, lengths, bert_dur, noise, _ = cases.e2e_inputs(case)
    sds = oracle_sds("ljspeech")
    cfg = cases.MODEL_CFGS["ljspeech"]
    steps = [torch.randn(1, 1, 256, generator=torch.Generator().manual_seed(i)) for i in range(2)]
    kw = dict(diffusion_steps=3, rng=dict(step_noises=steps), skip_decoder=True, forced_durations=torch.full((1, 6), 2.0))
    with torch.no_grad():
        base = O.synthesize(sds, cfg, tokens, lengths, bert_dur, noise, **kw)
        t0 = O.synthesize(sds, cfg, tokens, lengths, bert_dur, noise, s_prev=torch.randn(1, 256), t=0.0, **kw)
        fix = O.synthesize(sds, cfg, tokens, lengths, bert_dur, noise, s_prev=base["s_pred"], t=0.7, **kw)
    assert torch.equal(t0["s_carry"], base["s_pred"]) and torch.equal(t0["F0"], base["F0"])
    assert float((fix["s_carry"] - base["s_pred"]).abs().max()) < 1e-6
    assert base["wav"] is None
    case = dict(model="libritts", B=1, N=6, seed=4)
    tokens, lengths, bert_dur, noise, ref_s = cases.e2e_inputs(case)
    with torch.no_grad():
        ms = O.synthesize(oracle_sds("libritts"), cases.MODEL_CFGS["libritts"], tokens, lengths, bert_dur, noise, ref_s=ref_s, **kw)
    assert torch.equal(ms["s_carry"], torch.cat([ms["ref"], ms["s"]], dim=-1))


def test_equal_length_batch_plan_covers_every_utterance_once_and_balances():
    from styletts2_b200.parallel import plan_equal_length_batches
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(20, 60, (200,), generator=g).tolist()
    for world in (1, 2, 8):
        plan = plan_equal_length_batches(lengths, world, max_batch=8)
        seen = sorted(i for r in plan for b in r for i in b)
        assert seen == list(range(200))
        for r in plan:
            for b in r:
                assert 1 <= len(b) <= 8 and len({lengths[i] for i in b}) == 1      # equal token count inside a batch
        loads = [sum(lengths[i] for b in r for i in b) for r in plan]
        assert max(loads) - min(loads) <= 8 * 60                                        # LPT: within one batch of each other
        assert plan == plan_equal_length_batches(lengths, world, max_batch=8)            # deterministic


def test_fp16_two_plane_split_recipe_is_fp32_accurate():
    """CPU emulation of the GEMM precision recipe (csrc/linear_tc.cu): x = h + l*2^-11 with h = fp16(x),
    l = fp16((x-h)*2^11); product = h*h' + (h*l' + l*h')*2^-11 (l*l' dropped).  Without the 2^11 pre-scaling the low
    plane of small operands is an fp16 subnormal and the error is ~100x larger; with it the dot-product error stays
    at the fp32 level for operands spread over six decades."""
    g = torch.Generator().manual_seed(0)
    K = 1024
    a = (torch.randn(64, K, generator=g) * 10.0 ** (torch.rand(1, K, generator=g) * 6 - 4)).double()
    w = (torch.randn(48, K, generator=g) * 10.0 ** (torch.rand(48, 1, generator=g) * 3 - 3)).double()

    def split(x, scale):
        h = x.float().half()
        l = ((x.float() - h.float()) * scale).half()
        return h.double(), l.double()

    exact = a @ w.t()
    bound = a.abs() @ w.abs().t()
    errs = {}
    for scale in (2048.0, 1.0):
        ah, al = split(a, scale)
        wh, wl = split(w, scale)
        approx = ah @ wh.t() + (ah @ wl.t() + al @ wh.t()) / scale
        errs[scale] = float(((approx - exact).abs() / bound).max())
    fp32 = float((((a.float() @ w.float().t()).double() - exact).abs() / bound).max())
    assert errs[2048.0] < 4e-7, errs            # ~2^-22 per product, random signs
    assert errs[1.0] > 20 * errs[2048.0], errs  # unscaled low plane loses its bits to fp16 subnormals
    assert errs[2048.0] < 10 * max(fp32, 6e-8)


def test_pinning_records_are_within_the_stated_bars():
    """The committed pinning records (oracle vs the UNMODIFIED reference, written by oracle/make_golden*.py and
    oracle/pin_sweep.py in the build container) stay within the bars DESIGN.md states."""
    import json
    pin = json.load(open(os.path.join(GOLD, "PINNING.json")))
    sweep = json.load(open(os.path.join(GOLD, "PINNING_SWEEP.json")))
    records = list(pin["cases"].values()) + list(sweep["cases"].values())
    assert len(records) >= 9
    for rec in records:
        d = rec["diffs"]
        assert d.get("pred_dur_mismatch", 0) == 0
        for k, v in d.items():
            if k != "pred_dur_mismatch":
                assert v <= 1e-5 * max(1.0, rec.get("scale", {}).get(k, 1.0)), (k, v)
    st = pin["style_libri"]
    assert st["log_mel_max_abs"] <= 1e-4 and st["ref_s_max_abs"] <= 1e-5
