"""Per-modality hyper-parameters, same keys and values as the reference `config.configs`
(config.py:28-137; checked against a snapshot of the reference dict in
tests/test_host_logic.py).

Keys: input_dim/output_dim/hidden_dims (INR), data_dim, pixel_sizes, patch,
hierarchical_patch_nums, patch_nums (data & patching), latent_dim, fourier_dim, paddings,
layerwise_scale_factors, upsample_factors (latent grid, Fourier features, upsampler),
bitrate_range, lowest_bitrate (beta controller budget window).
"""


def _modality(data_dim, pixel_sizes, *, in_dim=32, out_dim=3, fourier_dim=16, patch_nums=None, level2=None,
              level3=None, first_scale=4, bitrate_range=0.3, lowest_bitrate=0.1):
    patched = patch_nums is not None
    return {
        'input_dim': in_dim,
        'output_dim': out_dim,
        'hidden_dims': [32] * 3,
        'data_dim': data_dim,
        'pixel_sizes': list(pixel_sizes),
        'patch': patched,
        'hierarchical_patch_nums': {'level2': list(level2), 'level3': list(level3)} if patched else None,
        'patch_nums': list(patch_nums) if patched else None,
        'latent_dim': 128,
        'fourier_dim': fourier_dim,
        'paddings': [2, 1, 1],
        'layerwise_scale_factors': [first_scale, 2, 2],
        'upsample_factors': [24, 16, 16] if data_dim == 3 else [16] * data_dim,
        'bitrate_range': bitrate_range,
        'lowest_bitrate': lowest_bitrate,
    }


configs = {
    'cifar': _modality(2, [32, 32]),
    'kodak': _modality(2, [64, 64], patch_nums=[512 // 64, 768 // 64], level2=[4, 4], level3=[8, 12],
                       bitrate_range=0.1, lowest_bitrate=0.05),
    'audio': _modality(1, [800], out_dim=1, patch_nums=[48000 // 800], level2=[4], level3=[60]),
    'video': _modality(3, [24, 16, 16], in_dim=34, fourier_dim=18, patch_nums=[24 // 24, 128 // 16, 128 // 16],
                       level2=[1, 4, 4], level3=[1, 8, 8], first_scale=(6, 4, 4)),
    'protein': _modality(1, [96]),
}
