"""ctypes binding of libsard_b200.so (include/sard_b200.h).  There is no CPU fallback: importing
this module without the built library raises, and every call checks the C status code."""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SARD_LIB') or os.path.join(HERE, 'libsard_b200.so')      # SARD_LIB: a tuning build (tools/)

MAX_LAYERS, MAX_MLP, MAX_IL, ACTION_DIM, EXT_DIM = 8, 4, 4, 8, 64
ACT = {'none': 0, None: 0, 'relu': 1, 'tanh': 2, 'sigmoid': 3}
POST_NONE, POST_SIN = 0, 1
PHASE_GATHER, PHASE_COMPUTE, PHASE_ALL = 1, 2, 3
CTL_FLOATS, CTL_INTS, MAX_GROUPS = 64, 32, 8
ADAM_STICKY = 1 << 30        # SARD_ADAM_STICKY: zero_grad() keeps zero tensors (torch < 2.0 semantics)

vp, ci, cll, cf, cd = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double


class SardArena(C.Structure):
    _fields_ = [('T', ci), ('n_total', ci), ('e_total', ci), ('d_obs', ci),
                ('obs', vp), ('actions', vp), ('node_ptr', vp), ('stage', vp), ('body_ind', vp),
                ('in_ptr', vp), ('in_nbr', vp), ('out_ptr', vp), ('out_nbr', vp), ('rep_local', vp),
                ('ext', vp * 3), ('ext_dim', ci * 3),
                ('values', vp), ('returns', vp), ('advantages', vp), ('fixed_log_probs', vp), ('max_nodes', ci)]


class SardSel(C.Structure):
    _fields_ = [('ids', vp), ('cum', vp), ('B', ci), ('n', ci), ('node_base', ci)]


class SardGatherOut(C.Structure):
    _fields_ = [('lead', ci), ('tail', ci), ('nconst', ci), ('onehot', ci), ('consts', vp),
                ('x', vp), ('ldx', ci), ('nd_graph', vp), ('nd_arena', vp), ('body', vp), ('root_row', vp)]


class SardGemm(C.Structure):
    _fields_ = [('A', vp), ('lda', ci), ('a_rows', vp),
                ('B0', vp), ('ldb0', ci), ('B1', vp), ('ldb1', ci), ('b_split', ci),
                ('b_group_stride', cll), ('bias', vp), ('bias_group_stride', ci),
                ('Y', vp), ('ldy', ci), ('y_rows', vp), ('Ypre', vp), ('Dact', vp), ('ldd', ci),
                ('act', ci), ('post', ci), ('dact0', ci), ('dact1', ci), ('dsplit', ci),
                ('M', ci), ('N', ci), ('R', ci), ('tiles', vp), ('num_tiles', vp), ('max_tiles', ci),
                ('agg_cols', ci), ('agg_ptr', vp), ('agg_nbr', vp), ('agg_node', vp), ('agg_graph', vp), ('agg_cum', vp),
                ('agg_node_base', ci), ('agg_out', vp), ('ld_agg_out', ci), ('b_ksplit', ci),
                ('splitk_ws', vp), ('splitk_floats', cll)]


class SardWGrad(C.Structure):
    _fields_ = [('P', vp), ('ldp', ci), ('p_rows', vp), ('Q', vp), ('ldq', ci), ('q_rows', vp),
                ('M', ci), ('N1', ci), ('N2', ci),
                ('chunks', vp), ('num_chunks', vp), ('max_chunks', ci), ('chunk_rows', ci),
                ('grp_chunk_ptr', vp), ('n_groups', ci), ('slot_of_group', vp),
                ('dW0', vp), ('ldw0', ci), ('dW1', vp), ('ldw1', ci), ('w_split', ci), ('w_slot_stride', cll),
                ('db', vp), ('b_slot_stride', ci), ('partials', vp)]


class SardPolicyLoss(C.Structure):
    _fields_ = [('out', vp), ('ld_out', ci), ('pre', vp), ('log_std', vp), ('d', ci), ('act_col', ci),
                ('tie_c0', ci), ('tie_c1', ci),
                ('nd_arena', vp), ('cum', vp), ('node_base', ci), ('ids', vp), ('B', ci),
                ('actions', vp), ('rep_local', vp), ('advantages', vp), ('fixed_log_probs', vp),
                ('log_prob_out', vp), ('log_prob_run', vp), ('tied_out', vp), ('dout', vp), ('ld_dout', ci),
                ('gstat', vp), ('clip_eps', cf), ('inv_B', cf),
                ('dout_sorted_hl', vp), ('spos', vp), ('srow', vp), ('rows_cap', ci), ('pad_', ci)]


class SardAdamSeg(C.Structure):
    _fields_ = [('p', vp), ('m', vp), ('v', vp), ('g', vp), ('n', cll), ('group', ci), ('pad', ci)]


class SardDense(C.Structure):
    _fields_ = [('out', ci), ('in_', ci), ('act', ci), ('w', cll), ('b', cll)]


class SardTower(C.Structure):
    _fields_ = [('D', ci), ('L', ci), ('act', ci), ('has_bias', ci), ('h', ci * (MAX_LAYERS + 1)),
                ('in_w', cll), ('in_b', cll),
                ('wl', cll * MAX_LAYERS), ('bl', cll * MAX_LAYERS), ('wr', cll * MAX_LAYERS),
                ('mean', vp), ('var', vp), ('std', vp), ('inv', vp), ('count', vp)]


class SardValueNet(C.Structure):
    _fields_ = [('tower', SardTower), ('enc', (SardDense * 2) * 3), ('n_mlp', ci), ('mlp', SardDense * MAX_MLP),
                ('head', SardDense), ('nconst', ci), ('consts', vp), ('onehot', ci), ('P', vp), ('G', vp)]


class SardIndexLinear(C.Structure):
    _fields_ = [('out', ci), ('in_', ci), ('W', vp), ('b', vp), ('gW', cll), ('gb', cll)]


class SardPolicyStage(C.Structure):
    _fields_ = [('stage', ci), ('tower', SardTower), ('lead', ci), ('tail', ci),
                ('n_il', ci), ('il', SardIndexLinear * MAX_IL), ('il_act', ci),
                ('out_dim', ci), ('post', ci), ('log_std', cll), ('tie_c0', ci), ('tie_c1', ci),
                ('max_index', ci), ('n_live', ci), ('slot_of_index', vp), ('GT', vp)]


class SardPolicyNet(C.Structure):
    _fields_ = [('stage', SardPolicyStage * 3), ('enc', (SardDense * 2) * 3), ('nconst', ci), ('consts', vp),
                ('P', vp), ('G', vp)]


class SardJsmlpTc(C.Structure):
    _fields_ = [('n', ci), ('in_dim', ci), ('hid', ci), ('out_dim', ci), ('act', ci), ('post', ci), ('max_index', ci),
                ('n_slots', ci), ('xs', vp), ('ld_xs', ci), ('W', vp * 3), ('b', vp * 3), ('slot_of_index', vp),
                ('tiles', vp), ('num_tiles', vp), ('max_tiles', ci), ('srow', vp), ('H', vp * 2), ('ldh', ci),
                ('Hhl', vp * 2), ('out', vp), ('pre', vp), ('ldo', ci), ('ws', vp), ('ws_floats', cll), ('padded', ci)]


class SardJsmlpTcBwd(C.Structure):
    _fields_ = [('dout', vp), ('ld_dout', ci), ('dOs_hl', vp), ('dZhl', vp * 2), ('dxs', vp), ('ld_dxs', ci),
                ('chunks', vp), ('num_chunks', vp), ('max_chunks', ci), ('grp_chunk_ptr', vp), ('partials', vp),
                ('gW', vp * 3), ('gb', vp * 3)]


P = C.POINTER
_PROTOS = {
    'sard_abi_version': (ci, []),
    'sard_last_error_string': (C.c_char_p, []),
    'sard_struct_sizes': (ci, [P(ci), ci]),
    'sard_launch_count': (cll, []),
    'sard_set_pdl': (ci, [ci]),
    'sard_get_pdl': (ci, []),
    'sard_set_wgrad_overlap': (ci, [ci]),
    'sard_set_adam_single_launch': (ci, [ci]),
    'sard_get_wgrad_overlap': (ci, []),
    'sard_prof_enable': (ci, [C.c_uint, ci]),
    'sard_prof_collect': (ci, [P(cd), ci]),
    'sard_prof_timeline': (ci, [P(cd), ci]),
    'sard_prof_kernels_enable': (ci, [ci]),
    'sard_prof_kernels_collect': (ci, [C.c_char_p, ci, P(cd), ci]),
    'sard_build_csr': (ci, [vp, vp, vp, vp, ci, vp, vp, vp, vp, vp]),
    'sard_orbit_rep': (ci, [vp, vp, ci, vp, ci, vp, vp, vp]),
    'sard_gather_nodes': (ci, [P(SardArena), P(SardSel), P(SardGatherOut), vp]),
    'sard_sort_by_index': (ci, [vp, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'sard_sort_by_index_padded': (ci, [vp, ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'sard_norm_moments': (ci, [vp, ci, ci, ci, vp, vp, vp]),
    'sard_norm_update': (ci, [vp, ci, ci, vp, vp, vp, vp, vp, vp]),
    'sard_norm_prepare_eval': (ci, [vp, ci, vp, vp]),
    'sard_norm_apply': (ci, [vp, ci, ci, ci, vp, vp, vp, cf, vp]),
    'sard_csr_aggregate': (ci, [vp, ci, vp, ci, vp, ci, vp, ci, ci, ci, ci, vp, vp, vp, ci, vp, vp, vp]),
    'sard_concat_sorted': (ci, [vp, ci, ci, vp, ci, ci, vp, vp, ci, vp, ci, vp]),
    'sard_unsort_grad': (ci, [vp, ci, vp, vp, ci, ci, vp, ci, ci, ci, vp, ci, vp, ci, ci, vp, ci, vp]),
    'sard_set_gemm_engine': (ci, [ci]),
    'sard_get_gemm_engine': (ci, []),
    'sard_linear_fwd': (ci, [P(SardGemm), vp]),
    'sard_gemm_agg_supported': (ci, [ci, ci]),
    'sard_linear_bwd_data': (ci, [P(SardGemm), vp]),
    'sard_linear_bwd_weight': (ci, [P(SardWGrad), vp]),
    'sard_wgrad_chunk_rows': (ci, [ci, ci, ci]),
    'sard_gauss_loss': (ci, [P(SardPolicyLoss), vp]),
    'sard_cat_loss': (ci, [P(SardPolicyLoss), vp]),
    'sard_loss_reduce': (ci, [vp, ci, ci, vp, vp, ci, vp, vp]),
    'sard_value_loss': (ci, [vp, vp, ci, vp, cf, vp, vp, vp, vp]),
    'sard_value_head': (ci, [vp, ci, vp, vp, ci, ci, ci, ci, vp, vp, cf, vp, vp, vp, vp, vp, ci, vp]),
    'sard_gae': (ci, [vp, vp, vp, ci, cd, cd, vp, vp, vp, vp]),
    'sard_gae_f64': (ci, [vp, vp, vp, ci, cd, cd, vp, vp, vp, vp]),
    'sard_allreduce_p2p_pad_bytes': (ci, []),
    'sard_allreduce_p2p': (ci, [vp, vp, ci, ci, cll, C.c_uint, vp]),
    'sard_allgather_p2p_pad_bytes': (ci, []),
    'sard_allgather_p2p': (ci, [vp, vp, ci, ci, cll, vp, ci, C.c_uint, vp]),
    'sard_sqnorm': (ci, [vp, cll, vp, vp, vp]),
    'sard_adam_prepare': (ci, [vp, cf, cf, cf, vp, cf, ci, ci, ci, cd, cd, vp, vp, vp, vp]),
    'sard_adam_apply': (ci, [vp, ci, cll, cf, cf, cf, cf, ci, vp, vp, vp]),
    'sard_adam_update': (ci, [vp, cf, cf, cf, vp, cll, vp, cf, ci, ci, ci, cd, cd, vp, vp, vp, vp, ci, cll, cf, cf, vp]),
    'sard_jsmlp_tc_supported': (ci, [ci, ci, ci, ci, ci]),
    'sard_jsmlp_tc_workspace': (cll, [ci, ci, ci, ci]),
    'sard_jsmlp_tc_presplit': (ci, [P(SardJsmlpTc), vp]),
    'sard_jsmlp_tc_split_input': (ci, [P(SardJsmlpTc), vp]),
    'sard_jsmlp_tc_xs_hl': (vp, [P(SardJsmlpTc)]),
    'sard_jsmlp_tc_forward': (ci, [P(SardJsmlpTc), vp]),
    'sard_jsmlp_tc_bwd_supported': (ci, [ci, ci, ci, ci, ci]),
    'sard_jsmlp_tc_partials_floats': (cll, [ci]),
    'sard_jsmlp_tc_backward': (ci, [P(SardJsmlpTc), P(SardJsmlpTcBwd), vp]),
    'sard_jsmlp_tc_wgrad': (ci, [P(SardJsmlpTc), P(SardJsmlpTcBwd), vp]),
    'sard_concat_sorted_split': (ci, [vp, ci, ci, vp, ci, ci, vp, vp, ci, vp, vp]),
    'sard_tc_debug_read': (ci, [P(cll), ci]),
    'sard_value_workspace': (cll, [P(SardValueNet), ci, ci]),
    'sard_policy_workspace': (cll, [P(SardPolicyNet), ci, ci, ci]),
    'sard_value_run': (ci, [P(SardValueNet), P(SardArena), P(SardSel), ci, ci, cf, vp, vp, ci, vp, vp, cll, vp]),
    'sard_policy_run': (ci, [P(SardPolicyNet), ci, P(SardArena), P(SardSel), ci, ci, cf, cf, vp, vp, ci, vp, vp, vp,
                             vp, cll, vp]),
}
EXPORTS = tuple(_PROTOS)


class SardError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        try:                                   # build in-tree when a toolchain is present
            from . import build as _build
            _build.build()
        except Exception as e:                 # noqa: BLE001
            raise ImportError(f'libsard_b200.so is not built ({LIB_PATH}) and could not be built: {e}; '
                              'run `python -m sard_b200.build` -- there is no CPU fallback') from e
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)              # AttributeError here = header/library mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.sard_abi_version() != 1:
        raise ImportError('libsard_b200.so ABI version mismatch')
    return lib


ABI_STRUCTS = (SardArena, SardSel, SardGatherOut, SardGemm, SardWGrad, SardPolicyLoss, SardAdamSeg, SardDense,
               SardTower, SardValueNet, SardIndexLinear, SardPolicyStage, SardPolicyNet, SardJsmlpTc, SardJsmlpTcBwd)

lib = _load()


def check_struct_sizes():
    buf = (ci * 32)()
    n = lib.sard_struct_sizes(buf, 32)
    got = [C.sizeof(s) for s in ABI_STRUCTS]
    want = list(buf[:n])
    if got != want:
        raise ImportError(f'ctypes struct layout {got} != C layout {want}')


check_struct_sizes()
if os.environ.get('SARD_PDL') is not None:               # 1 = programmatic dependent launch (default), 0 = off
    lib.sard_set_pdl(int(os.environ['SARD_PDL']))
if os.environ.get('SARD_WGRAD_OVERLAP') is not None:     # 1 = weight gradients on a side stream (default), 0 = off
    lib.sard_set_wgrad_overlap(int(os.environ['SARD_WGRAD_OVERLAP']))
if os.environ.get('SARD_GEMM_ENGINE') is not None:       # 0 = fp32 FFMA, 1 = tcgen05 3xTF32
    lib.sard_set_gemm_engine(int(os.environ['SARD_GEMM_ENGINE']))


def check(rc: int):
    if rc != 0:
        raise SardError(lib.sard_last_error_string().decode())


def ptr(t) -> vp:
    """Device pointer of a tensor (None -> NULL).  The caller keeps the tensor alive."""
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str = 'tensor'):
    if not t.is_cuda:
        raise SardError(f'{what} must live on a CUDA device: the SARD B200 path has no CPU implementation')
