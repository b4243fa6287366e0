"""tiny-cuda-nn-shaped modules (drop-in for ``import tinycudann as tcnn`` as used by
models/network_utils.py:47,90,181,209 and models/utils.py:119 of the reference).

Same constructor/config surface, same attributes (``n_input_dims``, ``n_output_dims``, one flat fp32
``params`` Parameter laid out network-first-then-grid so reference checkpoints load), same forward
contract (fp32 CUDA input in [0,1] -> fp16 output, differentiable w.r.t. params and -- for the hash
grid -- the input, including double backward).  Everything below is host logic; the arithmetic is in
libnsr_b200.so.  No CPU path: CPU tensors raise.
"""
import torch
import torch.nn as nn

from . import ops
from .lib import check_cuda


def free_temporary_memory():
    """tcnn.free_temporary_memory(): our kernels own no arena (the caller's torch allocator owns all
    memory), so this is a no-op kept for API compatibility (models/utils.py:119)."""
    return None


class _Module(nn.Module):
    """Shared parameter handling: flat fp32 master + cached fp16 copy refreshed when the master
    changes (tcnn re-casts the whole vector every forward: 75 MB of HBM traffic per call for the
    12.6 M-entry table; we only re-cast after an optimizer step)."""

    def __init__(self, seed=1337):
        super().__init__()
        self.seed = seed
        self.loss_scale = ops.LOSS_SCALE
        self._half_cache = None
        self._half_key = None

    def _params_half(self):
        p = self.params
        key = (p._version, p.data_ptr(), p.device)
        if self._half_key != key:
            self._half_cache = p.detach().to(torch.float16)
            self._half_key = key
        return self._half_cache

    def _check(self, x):
        check_cuda(x, self.params, what=type(self).__name__)
        if x.dim() != 2 or x.shape[1] != self.n_input_dims:
            raise RuntimeError(f'{type(self).__name__}: expected input [N,{self.n_input_dims}], got {tuple(x.shape)}')


class Encoding(_Module):
    """tcnn.Encoding(n_input_dims, encoding_config).  otype: HashGrid | SphericalHarmonics."""

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__(seed)
        self.n_input_dims = n_input_dims
        self.encoding_config = dict(encoding_config)
        self.otype = self.encoding_config.get('otype', 'HashGrid')
        g = torch.Generator().manual_seed(seed)
        if self.otype in ('HashGrid', 'Grid'):
            self.grid = ops.GridSpec(self.encoding_config, n_input_dims)
            self.n_output_dims = self.grid.n_output_dims
            init = (torch.rand(self.grid.n_params, generator=g) * 2 - 1) * 1e-4
        elif self.otype == 'SphericalHarmonics':
            if int(self.encoding_config.get('degree', 4)) != 4 or n_input_dims != 3:
                raise NotImplementedError('SphericalHarmonics: only degree 4 on 3-D inputs is implemented')
            self.grid = None
            self.n_output_dims = 16
            init = torch.zeros(0)
        else:
            raise NotImplementedError(f'encoding otype={self.otype!r} not implemented')
        self.params = nn.Parameter(init.float())
        self.dtype = dtype or torch.float16

    def forward(self, x):
        self._check(x)
        if self.grid is None:
            out = ops.sh4(x)
        else:
            out = ops.hashgrid(self.grid, x, self.params, self._params_half())
        return out if self.dtype == torch.float16 else out.to(self.dtype)


class Network(_Module):
    """tcnn.Network(n_input_dims, n_output_dims, network_config): FullyFusedMLP, 64 wide."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__(seed)
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.network_config = dict(network_config)
        self.mlp = ops.MlpSpec(n_input_dims, n_output_dims, self.network_config)
        g = torch.Generator().manual_seed(seed)
        self.params = nn.Parameter(self.mlp.init_params(g))

    def forward(self, x):
        self._check(x)
        return ops.mlp(self.mlp, x, self.params, self._params_half())


class NetworkWithInputEncoding(_Module):
    """tcnn.NetworkWithInputEncoding(n_input_dims, n_output_dims, encoding_config, network_config)
    (models/network_utils.py:209-214).  Flat params: MLP matrices first, then the grid levels."""

    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__(seed)
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.encoding_config, self.network_config = dict(encoding_config), dict(network_config)
        self.grid = ops.GridSpec(self.encoding_config, n_input_dims)
        self.mlp = ops.MlpSpec(self.grid.n_output_dims, n_output_dims, self.network_config)
        g = torch.Generator().manual_seed(seed)
        mlp_init = self.mlp.init_params(g)
        grid_init = (torch.rand(self.grid.n_params, generator=g) * 2 - 1) * 1e-4
        self.params = nn.Parameter(torch.cat([mlp_init, grid_init]).float())

    @property
    def n_mlp_params(self):
        return self.mlp.n_params

    def forward(self, x):
        self._check(x)
        ph = self._params_half()
        nm = self.mlp.n_params
        enc = ops.hashgrid(self.grid, x, self.params[nm:], ph[nm:])
        return ops.mlp(self.mlp, enc, self.params[:nm], ph[:nm])
