"""Model hyper-parameters of the two reference configurations (values of `model_params` in
Configs/config.yml:33-82 and Configs/config_libritts.yml; training-only keys omitted)."""

# model_params of Configs/config.yml:33-82 and Configs/config_libritts.yml (values only;
# the training-only keys are omitted).
MODEL_CFGS = {
    "ljspeech": dict(
        multispeaker=False, dim_in=64, hidden_dim=512, max_conv_dim=512, n_layer=3, n_mels=80,
        n_token=178, max_dur=50, style_dim=128, dropout=0.2,
        decoder=dict(type="istftnet", resblock_kernel_sizes=[3, 7, 11], upsample_rates=[10, 6],
                     upsample_initial_channel=512, resblock_dilation_sizes=[[1, 3, 5]] * 3,
                     upsample_kernel_sizes=[20, 12], gen_istft_n_fft=20, gen_istft_hop_size=5),
        diffusion=dict(embedding_mask_proba=0.1,
                       transformer=dict(num_layers=3, num_heads=8, head_features=64, multiplier=2),
                       dist=dict(sigma_data=0.2, estimate_sigma_data=True, mean=-3.0, std=1.0)),
        slm=dict(hidden=768, nlayers=13, initial_channel=64),
    ),
    "libritts": dict(
        multispeaker=True, dim_in=64, hidden_dim=512, max_conv_dim=512, n_layer=3, n_mels=80,
        n_token=178, max_dur=50, style_dim=128, dropout=0.2,
        decoder=dict(type="hifigan", resblock_kernel_sizes=[3, 7, 11], upsample_rates=[10, 5, 3, 2],
                     upsample_initial_channel=512, resblock_dilation_sizes=[[1, 3, 5]] * 3,
                     upsample_kernel_sizes=[20, 10, 6, 4]),
        diffusion=dict(embedding_mask_proba=0.1,
                       transformer=dict(num_layers=3, num_heads=8, head_features=64, multiplier=2),
                       dist=dict(sigma_data=0.2, estimate_sigma_data=True, mean=-3.0, std=1.0)),
        slm=dict(hidden=768, nlayers=13, initial_channel=64),
    ),
}
