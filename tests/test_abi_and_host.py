"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/nsr_b200.h declares
(no compute calls without a GPU), and the host logic (descriptors, bit packing, error behaviour)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'nsr_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nsr_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import nsr_b200
    path = nsr_b200.library_path()
    assert os.path.exists(path), 'build the library first: python instant-nsr-pl_b200/build.py'
    dll = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(dll, s), f'{s} declared in include/nsr_b200.h but not exported'
    dll.nsr_version.restype = ctypes.c_int
    assert dll.nsr_version() >= 100
    dll.nsr_last_error.restype = ctypes.c_char_p
    assert isinstance(dll.nsr_last_error(), bytes)
    # every symbol the python binding uses is declared in the header
    for s in nsr_b200.lib.symbols():
        assert s in syms, f'{s} bound in lib.py but missing from the header'


def test_struct_layouts_match_header():
    from nsr_b200.lib import GridT, MlpT, MarchT
    assert ctypes.sizeof(GridT) == 8 + 4 * 32 * 4 + 4
    assert ctypes.sizeof(MlpT) == 20
    assert ctypes.sizeof(MarchT) == 24 + 16


def test_grid_spec_matches_oracle_level_table():
    from nsr_b200.ops import GridSpec
    from oracle import hashgrid
    for cfg in [dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=1.447269237440378),
                dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=1.3195079107728942),
                dict(n_levels=8, n_features_per_level=2, log2_hashmap_size=12, base_resolution=4, per_level_scale=1.6)]:
        lt, gs = hashgrid.level_table(cfg), GridSpec(dict(cfg, otype='HashGrid'))
        assert np.array_equal(lt['scale'], gs.scale) and np.array_equal(lt['res'], gs.res)
        assert np.array_equal(lt['size'], gs.size) and np.array_equal(lt['offset'], gs.offset) and np.array_equal(lt['dense'], gs.dense)
        assert gs.n_params == lt['n_params']
        for l in range(gs.n_levels):
            assert gs.struct.scale[l] == lt['scale'][l] and gs.struct.offset[l] == lt['offset'][l]


def test_mlp_spec_and_unsupported_configs():
    from nsr_b200.ops import MlpSpec, GridSpec
    from oracle import mlp
    for n_in, n_out, nh in [(32, 16, 1), (32, 3, 2), (35, 13, 1)]:
        s = MlpSpec(n_in, n_out, dict(otype='FullyFusedMLP', activation='ReLU', output_activation='None', n_neurons=64, n_hidden_layers=nh))
        assert s.n_params == mlp.ffmlp_layout(n_in, n_out, 64, nh)[1]
    with pytest.raises(NotImplementedError):
        MlpSpec(32, 3, dict(otype='FullyFusedMLP', n_neurons=128, n_hidden_layers=2))
    with pytest.raises(NotImplementedError):
        GridSpec(dict(otype='HashGrid', n_levels=16, n_features_per_level=4))


def test_pack_binary_matches_oracle():
    from nsr_b200.nerfacc import pack_binary
    from oracle import occgrid
    b = torch.from_numpy(np.random.default_rng(0).random((16, 16, 16)) < 0.3)
    assert np.array_equal(pack_binary(b).numpy().view(np.uint32), occgrid.pack_bits(b.numpy()))


def test_cpu_tensors_are_rejected_loudly():
    from nsr_b200 import tcnn, nerfacc
    enc = tcnn.Encoding(3, dict(otype='HashGrid', n_levels=4, n_features_per_level=2, log2_hashmap_size=8, base_resolution=4,
                                per_level_scale=1.5))
    with pytest.raises(NotImplementedError):
        enc(torch.rand(4, 3))
    with pytest.raises(NotImplementedError):
        nerfacc.ray_marching(torch.zeros(2, 3), torch.ones(2, 3), render_step_size=0.1)
    with pytest.raises(NotImplementedError):
        nerfacc.render_weight_from_alpha(torch.rand(4, 1), ray_indices=torch.zeros(4, dtype=torch.long), n_rays=1)
    net = tcnn.NetworkWithInputEncoding(3, 16, dict(otype='HashGrid', n_levels=4, n_features_per_level=2, log2_hashmap_size=8,
                                                    base_resolution=4, per_level_scale=1.5),
                                        dict(otype='FullyFusedMLP', activation='ReLU', output_activation='None', n_neurons=64,
                                             n_hidden_layers=1))
    assert list(net.parameters())[0].dtype == torch.float32 and net.n_input_dims == 3 and net.n_output_dims == 16
    assert list(net.state_dict()) == ['params']


def test_fused_adamw_and_static_neus_refuse_cpu():
    """no CPU fallbacks: the optimizer and the static-shape NeuS path raise on CPU tensors instead of computing"""
    import torch
    from nsr_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    opt = FusedAdamW([p], lr=1e-2)
    with pytest.raises(NotImplementedError):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdamW([p], lr=-1.0)
    from nsr_b200 import models, configs
    m = models.make('neus', configs.neus_blender())
    m.train()
    m.background_color = torch.ones(3)
    with pytest.raises(NotImplementedError):
        m.forward_(torch.zeros(4, 6), static=True)
