"""build_model(...) with the reference's signature and container keys (models.py:614-694).

The inference hot path (SURVEY.md section 8 a-e) and the reference-style encoders (row f2) are built from kernels;
the training-only members of the reference's Munch (discriminators, aligner, pitch extractor) are outside the
scope table and are returned as nn.Identity() placeholders so notebook code that iterates over the
container (`[model[k].eval() for k in model]`) keeps working.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .diffusion import AudioDiffusionConditional, KDiffusion, LogNormalDistribution, StyleTransformer1d, Transformer1d
from .modules import Decoder, Linear, ProsodyPredictor, TextEncoder
from .synthetic import keyed_state_dict

HOT_MODULES = ["bert_encoder", "predictor", "decoder", "text_encoder", "diffusion"]
STYLE_MODULES = ["style_encoder", "predictor_encoder"]      # compute_style


class Munch(dict):
    """attribute dict (the reference returns munch.Munch)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def recursive_munch(d):
    """utils.py:63-69"""
    if isinstance(d, dict):
        return Munch((k, recursive_munch(v)) for k, v in d.items())
    if isinstance(d, list):
        return [recursive_munch(v) for v in d]
    return d


class _BertConfig:
    hidden_size = 768
    max_position_embeddings = 512


class BertStandIn(nn.Module):
    """PL-BERT is an input producer for this path: build_model only needs its config."""
    config = _BertConfig()

    def forward(self, *a, **k):
        raise RuntimeError("PL-BERT is not part of the accelerated path; pass bert_dur explicitly")


def build_model(args, text_aligner=None, pitch_extractor=None, bert=None):
    assert args.decoder.type in ["istftnet", "hifigan"], "Decoder type unknown"
    bert = bert if bert is not None else BertStandIn()
    dec = args.decoder
    kw = dict(dim_in=args.hidden_dim, style_dim=args.style_dim, dim_out=args.n_mels,
              resblock_kernel_sizes=dec.resblock_kernel_sizes, upsample_rates=dec.upsample_rates,
              upsample_initial_channel=dec.upsample_initial_channel, resblock_dilation_sizes=dec.resblock_dilation_sizes,
              upsample_kernel_sizes=dec.upsample_kernel_sizes)
    if dec.type == "istftnet":
        kw.update(gen_istft_n_fft=dec.gen_istft_n_fft, gen_istft_hop_size=dec.gen_istft_hop_size)
    decoder = Decoder(**kw)
    text_encoder = TextEncoder(channels=args.hidden_dim, kernel_size=5, depth=args.n_layer, n_symbols=args.n_token)
    predictor = ProsodyPredictor(style_dim=args.style_dim, d_hid=args.hidden_dim, nlayers=args.n_layer, max_dur=args.max_dur,
                                 dropout=args.dropout)
    tkw = dict(channels=args.style_dim * 2, context_embedding_features=bert.config.hidden_size, **args.diffusion.transformer)
    if args.multispeaker:
        transformer = StyleTransformer1d(context_features=args.style_dim * 2, **tkw)
    else:
        transformer = Transformer1d(**tkw)
    diffusion = AudioDiffusionConditional()
    diffusion.diffusion = KDiffusion(net=transformer,
                                     sigma_distribution=LogNormalDistribution(mean=args.diffusion.dist.mean, std=args.diffusion.dist.std),
                                     sigma_data=args.diffusion.dist.sigma_data, dynamic_threshold=0.0)
    diffusion.diffusion.net = transformer
    diffusion.unet = transformer
    ident = nn.Identity
    from .style import StyleEncoder   # models.py:639-640
    style_encoder = StyleEncoder(dim_in=args.dim_in, style_dim=args.style_dim, max_conv_dim=args.hidden_dim)
    predictor_encoder = StyleEncoder(dim_in=args.dim_in, style_dim=args.style_dim, max_conv_dim=args.hidden_dim)
    return Munch(bert=bert, bert_encoder=Linear(bert.config.hidden_size, args.hidden_dim), predictor=predictor, decoder=decoder,
                 text_encoder=text_encoder, predictor_encoder=predictor_encoder, style_encoder=style_encoder, diffusion=diffusion,
                 text_aligner=text_aligner if text_aligner is not None else ident(),
                 pitch_extractor=pitch_extractor if pitch_extractor is not None else ident(), mpd=ident(), msd=ident(), wd=ident())


def load_keyed_weights(model, seed: int = 0, voiced: bool = True, modules=None):
    """Deterministic synthetic weights (styletts2_b200/synthetic.py) into every hot-path module
    (modules=HOT_MODULES + STYLE_MODULES also fills the reference-style encoders)."""
    sds = {}
    for k in (modules or HOT_MODULES):
        shapes = {n: tuple(v.shape) for n, v in model[k].state_dict().items()}
        sd = keyed_state_dict(shapes, k, seed=seed, voiced=voiced)
        model[k].load_state_dict(sd)
        sds[k] = sd
    return sds


def load_checkpoint_params(model, params: dict):
    """Notebook cell 12: accept {'module_name': state_dict}, stripping a 'module.' prefix if present."""
    for key in model:
        if key in params and isinstance(model[key], nn.Module) and len(list(model[key].parameters())):
            sd = params[key]
            if any(k.startswith("module.") for k in sd):
                sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}
            model[key].load_state_dict(sd, strict=False)
    return model
