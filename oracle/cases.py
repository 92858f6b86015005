"""Golden-case definitions shared by oracle/make_golden.py and tests/ (TEST INFRASTRUCTURE)."""
import torch

_BS_SMALL = dict(dim=64, depth=2, stereo=True, num_stems=1, time_transformer_depth=1,
                 freq_transformer_depth=1, dim_head=64, heads=2, stft_n_fft=2048, stft_hop_length=441,
                 stft_win_length=2048, mask_estimator_depth=2, mlp_expansion_factor=2, flash_attn=True)

CASES = {
    # BS-RoFormer, default 62 bands, stereo, 1 stem, hop 441 (the BASELINE C2 shape scaled down)
    'bs_small': dict(kind='bs_roformer', seed=11, batch=2, length=441 * 40, cfg=dict(_BS_SMALL)),
    # mono, 2 stems, hop 512, deeper sub-transformers, skip connections, custom band table
    'bs_mono2': dict(kind='bs_roformer', seed=12, batch=1, length=512 * 24 + 100, cfg=dict(
        dim=32, depth=2, stereo=False, num_stems=2, time_transformer_depth=2, freq_transformer_depth=1,
        dim_head=64, heads=1, stft_n_fft=2048, stft_hop_length=512, stft_win_length=2048,
        mask_estimator_depth=3, mlp_expansion_factor=2, skip_connection=True,
        freqs_per_bands=[5] * 41 + [20] * 41)),
    # Mel-Band-RoFormer 60 mel bands, stereo, 2 stems, mask depth 2 (=> 3 Linears)
    'mel_small': dict(kind='mel_band_roformer', seed=13, batch=2, length=441 * 32, cfg=dict(
        dim=64, depth=1, stereo=True, num_stems=2, time_transformer_depth=1, freq_transformer_depth=1,
        num_bands=60, dim_head=64, heads=2, stft_n_fft=2048, stft_hop_length=441, stft_win_length=2048,
        mask_estimator_depth=2, mlp_expansion_factor=2, sample_rate=44100)),
    # Mel-Band-RoFormer with match_input_audio_length=True on a length the hop does not divide
    # (istft length = input length, mel_band_roformer.py:505,622-623), mono, 1 stem, 24 mel bands, hop 512
    'mel_match': dict(kind='mel_band_roformer', seed=15, batch=1, length=512 * 20 + 333, cfg=dict(
        dim=32, depth=1, stereo=False, num_stems=1, time_transformer_depth=1, freq_transformer_depth=1,
        num_bands=24, dim_head=64, heads=1, stft_n_fft=2048, stft_hop_length=512, stft_win_length=2048,
        mask_estimator_depth=1, mlp_expansion_factor=2, sample_rate=44100, match_input_audio_length=True)),
    # MDX23C TFC-TDF-v3, 2 instruments
    'mdx_small': dict(kind='mdx23c', seed=14, batch=2, length=256 * 31, cfg=dict(
        audio=dict(chunk_size=256 * 31, n_fft=1024, hop_length=256, dim_f=512, num_channels=2,
                   sample_rate=44100),
        model=dict(act='gelu', bottleneck_factor=4, growth=8, norm='InstanceNorm',
                   num_blocks_per_scale=2, num_channels=16, num_scales=2, num_subbands=4, scale=[2, 2]),
        training=dict(instruments=['vocals', 'other'], target_instrument=None, use_amp=False),
        inference=dict(batch_size=1, num_overlap=4))),
}

# (length, chunk_size, num_overlap, batch_size): short (< L/2), un-padded (< 2*border), long, ragged tails
DEMIX_IDENTITY_CASES = [
    (300, 1000, 4, 1), (300, 1000, 4, 2), (1400, 1000, 4, 2), (1501, 1000, 4, 2), (5003, 1000, 4, 1),
    (5003, 1000, 4, 2), (5003, 1000, 4, 4), (7777, 1000, 2, 2), (7777, 1000, 2, 3), (4000, 1000, 1, 1),
    (4321, 1000, 1, 2), (2600, 1001, 3, 2), (9999, 1000, 8, 4), (1000, 1000, 2, 1), (2001, 1000, 2, 1),
]

DEMIX_MODEL_CASES = {
    'demix_bs_ov2_b1': dict(model='bs_small', length=441 * 40 * 3 + 777, chunk_size=441 * 40, num_overlap=2,
                            batch_size=1, instruments=['vocals', 'other'], target='vocals', seed=21),
    'demix_bs_ov4_b2': dict(model='bs_small', length=441 * 40 * 2 + 5000, chunk_size=441 * 40, num_overlap=4,
                            batch_size=2, instruments=['vocals', 'other'], target='vocals', seed=22),
    'demix_mel_ov2_b2': dict(model='mel_small', length=441 * 32 * 2 + 333, chunk_size=441 * 32, num_overlap=2,
                             batch_size=2, instruments=['vocals', 'other'], target=None, seed=23),
}


# Per-file flows of inference_pytorch.run_folder_pytorch_optimized (:219-260) run by the UNMODIFIED reference:
# normalize -> demix -> TTA -> DemudPhaseRemix -> instrumental -> denormalize.  'flow_bs_*' takes the branch without an
# 'instrumental' stem (:236-242), 'flow_mel_*' the branch with one (:243-250, where apply_tta updates the first pass's
# estimates in place).
FLOW_CASES = {
    'flow_bs_norm_tta_demud': dict(model='bs_small', length=441 * 40 * 2 + 900, chunk_size=441 * 40, num_overlap=2,
                                   batch_size=2, instruments=['vocals', 'other'], target='vocals', seed=81,
                                   normalize=True, use_tta=True, demud=True, extract_instrumental=True),
    'flow_mel_tta_demud_inst': dict(model='mel_small', length=441 * 32 * 2 + 100, chunk_size=441 * 32, num_overlap=2,
                                    batch_size=1, instruments=['vocals', 'instrumental'], target=None, seed=82,
                                    normalize=False, use_tta=True, demud=True, extract_instrumental=False),
    'flow_bs_norm_only': dict(model='bs_small', length=441 * 40 + 4321, chunk_size=441 * 40, num_overlap=4,
                              batch_size=1, instruments=['vocals', 'other'], target='vocals', seed=83,
                              normalize=True, use_tta=False, demud=False, extract_instrumental=False),
}


def make_input(case):
    c = 1 if case['kind'] != 'mdx23c' and not case['cfg'].get('stereo', False) else 2
    g = torch.Generator().manual_seed(case['seed'] + 1000)
    return 0.5 * torch.randn(case['batch'], c, case['length'], generator=g)
