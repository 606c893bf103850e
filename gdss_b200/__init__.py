"""gdss_b200 -- B200-native (sm_100a) implementation of the GDSS reverse-diffusion sampling hot path.

Host-side mirror of the reference's Python interface for that path (same names, arguments and
error behaviour); all arithmetic on tensors is done by hand-written CUDA kernels behind the C ABI
declared in include/gdss_b200.h.

    from gdss_b200.solver import get_pc_sampler, S4_solver          # solver.py:153, :200
    from gdss_b200.sde import VPSDE, VESDE, subVPSDE                # sde.py
    from gdss_b200.models import ScoreNetworkX, ScoreNetworkX_GMH, ScoreNetworkA
    from gdss_b200.graph_utils import mask_x, mask_adjs, gen_noise, pow_tensor, quantize, quantize_mol
"""
__version__ = "0.1.0"
