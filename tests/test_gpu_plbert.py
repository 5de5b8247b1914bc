"""GPU tier: PL-BERT on our kernels vs its reference implementation, transformers.AlbertModel
(the third-party dependency the reference wraps, Utils/PLBERT/util.py:4-12; installed version is the de-facto pin)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import maxdiff, record

D = "cuda:0"


def _models():
    from transformers import AlbertConfig, AlbertModel
    from styletts2_b200.plbert import PLBert
    from styletts2_b200.synthetic import keyed_state_dict
    cfg = AlbertConfig(vocab_size=178, hidden_size=768, num_attention_heads=12, intermediate_size=2048, max_position_embeddings=512,
                       num_hidden_layers=12)
    ref = AlbertModel(cfg).eval()
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items() if v.dtype == torch.float32}
    sd = keyed_state_dict(shapes, "plbert")
    ref.load_state_dict(sd, strict=False)
    ours = PLBert().to(D).eval()
    missing, unexpected = ours.load_state_dict(sd, strict=False)
    assert not missing, missing
    return ref, ours


@pytest.mark.parametrize("B,N,ragged", [(2, 23, False), (3, 40, True), (4, 128, False)])
def test_plbert_matches_transformers_albert(B, N, ragged):
    ref, ours = _models()
    g = torch.Generator().manual_seed(B * 100 + N)
    tokens = torch.randint(0, 178, (B, N), generator=g)
    lengths = torch.tensor([N] + [max(3, N - 7 * (i + 1)) for i in range(B - 1)]) if ragged else torch.full((B,), N)
    mask = (torch.arange(N)[None] < lengths[:, None]).int()
    with torch.no_grad():
        want = ref(tokens, attention_mask=mask).last_hidden_state
        got = ours(tokens.to(D), attention_mask=mask.to(D)).cpu()
    # padded query rows are meaningless in both (the glue masks them downstream): compare valid rows
    valid = mask.bool()
    d = float((got - want)[valid].abs().max())
    scale = float(want[valid].abs().max())
    record("plbert", B=B, N=N, ragged=ragged, maxabs=d, scale=scale)
    assert d <= 2e-4 * max(1.0, scale), (d, scale)


def test_plbert_state_dict_matches_albert_checkpoint_schema():
    from transformers import AlbertConfig, AlbertModel
    from styletts2_b200.plbert import PLBert
    cfg = AlbertConfig(vocab_size=178, hidden_size=768, num_attention_heads=12, intermediate_size=2048, max_position_embeddings=512,
                       num_hidden_layers=12)
    want = {k: tuple(v.shape) for k, v in AlbertModel(cfg).state_dict().items() if v.dtype == torch.float32}
    have = {k: tuple(v.shape) for k, v in PLBert().state_dict().items()}
    assert want == have
