"""Host-side mirror of the reference's PnP op interface over the CUDA C ABI.

Same names, argument meaning and return tuples as ``monorun/ops/least_squares``:

* :func:`pnp_uncert`   <- monorun/ops/least_squares/pnp_uncert.py:7-87
* :class:`PnPUncert`   <- monorun/ops/least_squares/pnp_uncert.py:90-142 (registered as ``'PnPUncert'`` in ``PNP``)
* :func:`build_pnp`    <- monorun/ops/least_squares/builder.py:6-7

What changes underneath: no device->host copy (pnp_uncert.py:34-43), no per-object Python loop
(pnp_uncert_cpu.py:180-191), no OpenCV / Ceres; one batched sm_100a kernel launch solves every object and also
produces the pose covariance that the reference computes with ~40 torch launches (pnp_uncert.py:71-85).
PyTorch is used for device memory and streams only.  There is no CPU fallback: tensors must live on a CUDA
device and ``libmonorun_pnp.so`` must be built, otherwise these calls raise.
"""
import torch

from . import _native
from .registry import PNP, build_pnp  # noqa: F401  (re-exported like monorun.ops)

C = _native.CONST
RESULT_STRIDE = C['MRPNP_RESULT_STRIDE']

_PREC = {'fp64': C['MRPNP_PREC_FP64'], 'mixed': C['MRPNP_PREC_MIXED'], 'fast': C['MRPNP_PREC_FAST']}

_ctx_cache = {}


class _Ctx:
    """One ``mrpnp_ctx`` per CUDA device, created on first use."""

    def __init__(self, device_index):
        lib = _native.lib()
        out = _native.ffi.new('mrpnp_ctx**')
        _native.check(lib.mrpnp_create(out, device_index))
        self.ptr = out[0]
        self.device_index = device_index

    @property
    def launches(self):
        return int(_native.lib().mrpnp_launch_count(self.ptr))


def get_ctx(device):
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('monorun_b200 PnP runs on CUDA tensors only (no CPU fallback); got device '
                           f'{device}')
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx_cache:
        _ctx_cache[idx] = _Ctx(idx)
    return _ctx_cache[idx]


def launch_count(device='cuda'):
    return get_ctx(device).launches


def handed_back_count(device='cuda'):
    """Objects the fp32 fast path handed to the exact fp64 routine so far on this device (synchronises)."""
    return int(_native.lib().mrpnp_handed_back_count(get_ctx(device).ptr))


def make_params(n_obj, n_pts, **kw):
    p = _native.ffi.new('mrpnp_params*')
    _native.lib().mrpnp_default_params(p, n_obj, n_pts)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def pack_mask(mask):
    """[N,P] bool -> [N,ceil(P/32)] int32 words (bit j of word k = point 32k+j): the C ABI's mask format."""
    n, p = mask.shape
    w = (p + 31) // 32
    m = torch.zeros((n, w * 32), dtype=torch.int64, device=mask.device)
    m[:, :p] = mask.to(torch.int64)
    words = (m.view(n, w, 32) << torch.arange(32, device=mask.device, dtype=torch.int64)).sum(-1)
    return words.view(torch.int32)[:, ::2].contiguous()  # low 32 bits (little endian)


def unpack_mask(words, n_pts):
    """Inverse of :func:`pack_mask`."""
    n, w = words.shape
    bits = (words.to(torch.int64).unsqueeze(-1) >> torch.arange(32, device=words.device, dtype=torch.int64)) & 1
    return bits.view(n, w * 32)[:, :n_pts].bool()


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t, ctype='float*'):
    return _native.ffi.cast(ctype, t.data_ptr()) if t is not None else _native.ffi.NULL


def solve_batched(coords_3d, coords_2d, weights, cam_mats, uv_range, init_pose=None, inlier_mask=None, *,
                  layout='planar', weight_mode='logstd', z_min=0.5, std_scale=10.0, istd_thres=0.6,
                  inlier_opt_only=True, cov_mode='pipeline', precision='fast', max_iterations=50,
                  adopt_candidate_on_ftol=False, return_inlier_mask=True, return_fp64=False, peers=None, row_offset=0,
                  decision_bands=None, ransac_thres=None, peer_flags=None, flag_slot=0, flag_value=0, acks=None,
                  ack_value=0, hand_back_log=None):
    """Batched uncertainty-PnP on device tensors -- direct wrapper of ``mrpnp_solve``.

    layout 'planar':      coords_3d [N,3,*], coords_2d [N,2,*], weights [N,2|3,*]   (head level)
    layout 'interleaved': coords_3d [N,P,3], coords_2d [N,P,2], weights [N,P,2|3]   (op level)
    weight_mode 'logstd' | 'istd' | 'full';  cam_mats [N|1,3,3];  uv_range [N|1,4] = u_min,u_max,v_min,v_max;
    init_pose [N,4] or None (on-device linear initialiser);  inlier_mask [N,P] bool/uint8 or None.

    ransac_thres [N] float (pixels) or None: reprojection-threshold consensus after the start pose, the deterministic
    counterpart of the inlier refinement of cv2.solvePnPRansac (pnp_uncert_cpu.py:34-51; include/monorun_pnp.h).

    peers: device pointers (ints, valid on this device) of every rank's gathered [n_total,24] buffer -- the kernel then
    stores each result row into ALL of them at row ``row_offset + i`` (fused all-gather, see ``dist.FusedGather``) and
    the returned ``result`` is None.

    Returns (result [N,24] float32, inlier_mask [N,P] bool | None, result64 [N,8] float64 | None); the result
    row is ``yaw,tx,ty,tz | cov(16) | valid, lm_iterations, final_cost, trust_region_radius``.
    """
    dev = coords_3d.device
    ctx = get_ctx(dev)
    n = coords_3d.shape[0]
    planar = layout == 'planar'
    n_pts = coords_3d[0].numel() // 3 if n else (coords_3d.shape[2:].numel() if planar else coords_3d.shape[1])
    wmode = {'logstd': C['MRPNP_W_LOGSTD'], 'istd': C['MRPNP_W_ISTD'], 'full': C['MRPNP_W_FULL']}[weight_mode]
    result = torch.empty((n, RESULT_STRIDE), dtype=torch.float32, device=dev) if not peers else None
    n_words = (n_pts + 31) // 32
    inl_out = torch.empty((n, n_words), dtype=torch.int32, device=dev) if return_inlier_mask else None
    res64 = torch.empty((n, 8), dtype=torch.float64, device=dev) if return_fp64 else None
    if n == 0:
        return result, (unpack_mask(inl_out, n_pts) if inl_out is not None else None), res64
    c3, c2, w = _f32c(coords_3d), _f32c(coords_2d), _f32c(weights)
    cam, rng = _f32c(cam_mats).reshape(-1, 9), _f32c(uv_range).reshape(-1, 4)
    if cam.shape[0] not in (1, n) or rng.shape[0] not in (1, n):
        raise ValueError('cam_mats / uv_range must have batch size 1 or N')
    init = _f32c(init_pose) if init_pose is not None else None
    inl_in = pack_mask(inlier_mask.reshape(n, n_pts).bool()) if inlier_mask is not None else None
    p = make_params(
        n, n_pts,
        layout=C['MRPNP_LAYOUT_PLANAR'] if planar else C['MRPNP_LAYOUT_INTERLEAVED'],
        weight_mode=wmode,
        cam_stride=9 if cam.shape[0] == n and n > 1 else 0,
        range_stride=4 if rng.shape[0] == n and n > 1 else 0,
        precision=_PREC[precision],
        cov_mode={'none': 0, 'pipeline': 1, 'ceres': 2}[cov_mode],
        init_mode=C['MRPNP_INIT_GIVEN'] if init is not None else C['MRPNP_INIT_LINEAR'],
        inlier_opt_only=int(bool(inlier_opt_only)), max_iterations=int(max_iterations),
        adopt_candidate_on_ftol=int(bool(adopt_candidate_on_ftol)),
        z_min=float(z_min), std_scale=float(std_scale), istd_thres=float(istd_thres))
    thr = None
    if ransac_thres is not None:
        thr = _f32c(ransac_thres).reshape(-1)
        if thr.numel() != n or thr.device != dev:
            raise ValueError('ransac_thres must be a device tensor with one threshold per object')
        p.ransac_thres = _ptr(thr)
    if decision_bands is not None:   # (band_first, band_rel, band_mix) of mrpnp_params; the defaults are the product's
        p.band_first, p.band_rel, p.band_mix = (float(v) for v in decision_bands[:3])
        if len(decision_bands) > 3:   # (…, band_ratio, band_rel_min): the adaptive relative band
            p.band_ratio, p.band_rel_min = float(decision_bands[3]), float(decision_bands[4])
            p.band_ratio_from = int(decision_bands[5]) if len(decision_bands) > 5 else 0
    if hand_back_log is not None:   # int32 [N] device tensor: reason | evaluations << 8 of handed-back objects
        assert hand_back_log.dtype == torch.int32 and hand_back_log.numel() == n and hand_back_log.device == dev
        p.hand_back_log = _native.ffi.cast('int32_t*', hand_back_log.data_ptr())
    if peers:
        if len(peers) > C['MRPNP_MAX_PEERS']:
            raise ValueError('at most %d peers' % C['MRPNP_MAX_PEERS'])
        p.n_peers, p.row_offset = len(peers), int(row_offset)
        for r, ptr in enumerate(peers):
            p.peer_results[r] = _native.ffi.cast('float*', int(ptr))
        if peer_flags:   # completion flags raised by the launch's last thread block (dist.FusedGather)
            assert len(peer_flags) == len(peers)
            for r, ptr in enumerate(peer_flags):
                p.peer_flags[r] = _native.ffi.cast('uint32_t*', int(ptr))
            p.flag_slot, p.flag_value = int(flag_slot), int(flag_value) & 0xffffffff
            if acks:
                p.acks, p.ack_value = _native.ffi.cast('const uint32_t*', int(acks)), int(ack_value) & 0xffffffff
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_solve(
            ctx.ptr, p, _ptr(c3), _ptr(c2), _ptr(w), _ptr(cam), _ptr(rng), _ptr(init),
            _ptr(inl_in, 'uint32_t*'), _ptr(result), _ptr(inl_out, 'uint32_t*'), _ptr(res64, 'double*'),
            _native.ffi.cast('void*', stream)))
    return result, (unpack_mask(inl_out, n_pts) if inl_out is not None else None), res64


def solve_noc_batched(coords_3d, coords_2d, weights, logdim, logdim_wgt, cam_mats, uv_range, init_dimpose,
                      inlier_mask=None, *, layout='planar', weight_mode='istd', z_min=0.5, huber_delta=1.0,
                      max_iterations=50):
    """The reference's 7-parameter solvers, batched on the device -- direct wrapper of ``mrpnp_solve_noc``
    (``pnp_noc_uncert`` for weight_mode 'istd', ``pnp_noc_cov_uncert`` for 'full'; ext.h:15-43).

    coords_3d holds NORMALISED object coordinates; the unknowns are [log l, log h, log w, yaw, tx, ty, tz] with the
    prior ``logdim_wgt * (x[:3] - logdim)`` and ``HuberLoss(huber_delta)`` on every residual block.  Tensor shapes as
    :func:`solve_batched`; logdim, logdim_wgt [N,3]; init_dimpose [N,7].

    Returns result [N,12] float64: dimpose[7], valid, lm_iterations, final_cost, cost_evals, termination."""
    dev = coords_3d.device
    ctx = get_ctx(dev)
    n = coords_3d.shape[0]
    planar = layout == 'planar'
    n_pts = coords_3d[0].numel() // 3 if n else (coords_3d.shape[2:].numel() if planar else coords_3d.shape[1])
    if weight_mode not in ('istd', 'full'):
        raise ValueError("weight_mode must be 'istd' or 'full'")
    result = torch.empty((n, 12), dtype=torch.float64, device=dev)
    if n == 0:
        return result
    c3, c2, w = _f32c(coords_3d), _f32c(coords_2d), _f32c(weights)
    cam, rng = _f32c(cam_mats).reshape(-1, 9), _f32c(uv_range).reshape(-1, 4)
    if cam.shape[0] not in (1, n) or rng.shape[0] not in (1, n):
        raise ValueError('cam_mats / uv_range must have batch size 1 or N')
    ld, lw, init = _f32c(logdim).reshape(n, 3), _f32c(logdim_wgt).reshape(n, 3), _f32c(init_dimpose).reshape(n, 7)
    inl_in = pack_mask(inlier_mask.reshape(n, n_pts).bool()) if inlier_mask is not None else None
    p = _native.ffi.new('mrpnp_noc_params*')
    p.n_obj, p.n_pts = n, n_pts
    p.layout = C['MRPNP_LAYOUT_PLANAR'] if planar else C['MRPNP_LAYOUT_INTERLEAVED']
    p.weight_mode = C['MRPNP_W_FULL'] if weight_mode == 'full' else C['MRPNP_W_ISTD']
    p.cam_stride = 9 if cam.shape[0] == n and n > 1 else 0
    p.range_stride = 4 if rng.shape[0] == n and n > 1 else 0
    p.max_iterations = int(max_iterations)
    p.z_min, p.huber_delta = float(z_min), float(huber_delta)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_solve_noc(
            ctx.ptr, p, _ptr(c3), _ptr(c2), _ptr(w), _ptr(ld), _ptr(lw), _ptr(cam), _ptr(rng), _ptr(init),
            _ptr(inl_in, 'uint32_t*'), _ptr(result, 'double*'), _native.ffi.cast('void*', stream)))
    return result


def solve_6dof_batched(coords_3d, coords_2d, weights, cam_mats, uv_range, init_pose6, inlier_mask=None, *,
                       layout='planar', weight_mode='logstd', z_min=0.5, std_scale=10.0, max_iterations=50,
                       precision='mixed'):
    """6-DoF extension of the solver -- direct wrapper of ``mrpnp_solve_6dof``: unknowns [rvec(3), t(3)] (angle-axis
    as ceres::AngleAxisRotatePoint), everything else as :func:`solve_batched`.  init_pose6 [N,6].
    precision: 'mixed' (default; 'fast' is accepted as a synonym) -- correspondences staged once in shared memory, fp64
    cost chain, fp32 normal equations (pnp_6dof_fast.cuh); 'fp64' -- everything in fp64 (pnp_6dof.cuh).
    Returns result [N,48] float64: rvec, t | cov 6x6 | valid, lm_iterations, final_cost, cost_evals, termination, pad."""
    dev = coords_3d.device
    ctx = get_ctx(dev)
    n = coords_3d.shape[0]
    planar = layout == 'planar'
    n_pts = coords_3d[0].numel() // 3 if n else (coords_3d.shape[2:].numel() if planar else coords_3d.shape[1])
    wmode = {'logstd': C['MRPNP_W_LOGSTD'], 'istd': C['MRPNP_W_ISTD'], 'full': C['MRPNP_W_FULL']}[weight_mode]
    result = torch.empty((n, 48), dtype=torch.float64, device=dev)
    if n == 0:
        return result
    c3, c2, w = _f32c(coords_3d), _f32c(coords_2d), _f32c(weights)
    cam, rng = _f32c(cam_mats).reshape(-1, 9), _f32c(uv_range).reshape(-1, 4)
    if cam.shape[0] not in (1, n) or rng.shape[0] not in (1, n):
        raise ValueError('cam_mats / uv_range must have batch size 1 or N')
    init = _f32c(init_pose6).reshape(n, 6)
    inl_in = pack_mask(inlier_mask.reshape(n, n_pts).bool()) if inlier_mask is not None else None
    p = make_params(
        n, n_pts, layout=C['MRPNP_LAYOUT_PLANAR'] if planar else C['MRPNP_LAYOUT_INTERLEAVED'], weight_mode=wmode,
        cam_stride=9 if cam.shape[0] == n and n > 1 else 0, range_stride=4 if rng.shape[0] == n and n > 1 else 0,
        max_iterations=int(max_iterations), z_min=float(z_min), std_scale=float(std_scale),
        precision=_PREC[precision])
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_solve_6dof(
            ctx.ptr, p, _ptr(c3), _ptr(c2), _ptr(w), _ptr(cam), _ptr(rng), _ptr(init), _ptr(inl_in, 'uint32_t*'),
            _ptr(result, 'double*'), _native.ffi.cast('void*', stream)))
    return result


def exact_hessian(coords_3d, coords_2d, weights, cam_mats, uv_range, pose, inlier_mask=None, *, layout='planar',
                  weight_mode='logstd', z_min=0.5, std_scale=10.0, rows=None, return_hessian=True):
    """``mrpnp_exact_hessian``: the second-order pose Hessian of the reference's ``exact_hessian`` (hessian.py:5-64)
    at ``pose`` ([N,4] yaw,t -- or the [N,24] result rows of :func:`solve_batched`).  With ``rows`` ([N,24], may be
    the same tensor as ``pose``) the inverse is written into the rows' covariance slots and ``valid`` is cleared where
    the Hessian is singular.  Returns H [N,4,4] float32 (or None with return_hessian=False)."""
    dev = coords_3d.device
    ctx = get_ctx(dev)
    n = coords_3d.shape[0]
    planar = layout == 'planar'
    n_pts = coords_3d[0].numel() // 3 if n else (coords_3d.shape[2:].numel() if planar else coords_3d.shape[1])
    if weight_mode not in ('logstd', 'istd'):
        raise ValueError("the exact Hessian takes per-axis weights: weight_mode 'logstd' or 'istd'")
    h = torch.empty((n, 4, 4), dtype=torch.float32, device=dev) if return_hessian else None
    if n == 0:
        return h
    if pose.dtype != torch.float32 or not pose.is_contiguous() or pose.dim() != 2 or pose.shape[1] < 4:
        raise ValueError('pose must be a contiguous float32 [N, >=4] tensor')
    if rows is not None and (rows.dtype != torch.float32 or not rows.is_contiguous() or rows.shape != (n, RESULT_STRIDE)):
        raise ValueError('rows must be a contiguous float32 [N,%d] tensor' % RESULT_STRIDE)
    c3, c2, w = _f32c(coords_3d), _f32c(coords_2d), _f32c(weights)
    cam, rng = _f32c(cam_mats).reshape(-1, 9), _f32c(uv_range).reshape(-1, 4)
    if cam.shape[0] not in (1, n) or rng.shape[0] not in (1, n):
        raise ValueError('cam_mats / uv_range must have batch size 1 or N')
    inl_in = pack_mask(inlier_mask.reshape(n, n_pts).bool()) if inlier_mask is not None else None
    p = make_params(
        n, n_pts, layout=C['MRPNP_LAYOUT_PLANAR'] if planar else C['MRPNP_LAYOUT_INTERLEAVED'],
        weight_mode=C['MRPNP_W_LOGSTD'] if weight_mode == 'logstd' else C['MRPNP_W_ISTD'],
        cam_stride=9 if cam.shape[0] == n and n > 1 else 0, range_stride=4 if rng.shape[0] == n and n > 1 else 0,
        z_min=float(z_min), std_scale=float(std_scale))
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_exact_hessian(
            ctx.ptr, p, _ptr(c3), _ptr(c2), _ptr(w), _ptr(cam), _ptr(rng), _ptr(pose), pose.shape[1],
            _ptr(inl_in, 'uint32_t*'), _ptr(h), _ptr(rows), _native.ffi.cast('void*', stream)))
    return h


def pose_features(rows, dims, cov_calib_logscale=None, cov_correction_sd=0.0, distance_z_depth=False, use_calib=False,
                  pose_norm=None):
    """``mrpnp_pose_features``: result rows [N,24] + dims [N,3] -> (features [N,17], pose_cov_calib [N,16])."""
    dev = rows.device
    ctx = get_ctx(dev)
    n = rows.shape[0]
    rows, dims = _f32c(rows), _f32c(dims)
    feat = torch.empty((n, 17), dtype=torch.float32, device=dev)
    cal = torch.empty((n, 16), dtype=torch.float32, device=dev)
    if n == 0:
        return feat, cal
    ls = _f32c(cov_calib_logscale) if cov_calib_logscale is not None else None
    nm = [None] * 4
    eps = 0.0
    if pose_norm is not None:
        nm = [_f32c(t.detach()) for t in (pose_norm.running_mean, pose_norm.running_var, pose_norm.weight, pose_norm.bias)]
        eps = float(pose_norm.eps)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_pose_features(
            ctx.ptr, _ptr(rows), _ptr(dims), _ptr(ls), float(cov_correction_sd), int(bool(distance_z_depth)),
            int(bool(use_calib)), _ptr(nm[0]), _ptr(nm[1]), _ptr(nm[2]), _ptr(nm[3]), eps, _ptr(feat), _ptr(cal), n,
            _native.ffi.cast('void*', stream)))
    return feat, cal


def finish_scores(score_logits, rows, dims, det_scores=None, pre_sigmoid=True):
    """``mrpnp_finish_scores``: -> (scores [N], bbox_3d [N,8] = l,h,w,x,y,z,ry,score)."""
    dev = rows.device
    ctx = get_ctx(dev)
    n = rows.shape[0]
    scores = torch.empty((n,), dtype=torch.float32, device=dev)
    bbox = torch.empty((n, 8), dtype=torch.float32, device=dev)
    if n == 0:
        return scores, bbox
    lg, rows, dims = _f32c(score_logits), _f32c(rows), _f32c(dims)
    ds = _f32c(det_scores) if det_scores is not None else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_finish_scores(ctx.ptr, _ptr(lg), _ptr(rows), _ptr(dims), _ptr(ds),
                                                        int(bool(pre_sigmoid)), _ptr(scores), _ptr(bbox), n,
                                                        _native.ffi.cast('void*', stream)))
    return scores, bbox


def score_stage(rows, dims, reg_fc_out, w1, b1, w2t, b2, w3, b3, cov_calib_logscale=None, cov_correction_sd=0.0,
                distance_z_depth=False, use_calib=False, pose_norm=None, det_scores=None, pre_sigmoid=True,
                return_logits=False):
    """``mrpnp_score_stage``: result rows -> (scores [N], bbox_3d [N,8], pose_cov_calib [N,16], logits [N] | None) in
    ONE launch (features, the three Linear layers of MLPScoreHead, sigmoid / invalid / 2-D score, result rows)."""
    dev = rows.device
    ctx = get_ctx(dev)
    n = rows.shape[0]
    scores = torch.empty((n,), dtype=torch.float32, device=dev)
    bbox = torch.empty((n, 8), dtype=torch.float32, device=dev)
    cal = torch.empty((n, 16), dtype=torch.float32, device=dev)
    logits = torch.empty((n,), dtype=torch.float32, device=dev) if return_logits else None
    if n == 0:
        return scores, bbox, cal, logits
    rows, dims = _f32c(rows), _f32c(dims)
    reg = _f32c(reg_fc_out) if reg_fc_out is not None else None
    ls = _f32c(cov_calib_logscale) if cov_calib_logscale is not None else None
    nm, eps = [None] * 4, 0.0
    if pose_norm is not None:
        nm = [_f32c(t.detach()) for t in (pose_norm.running_mean, pose_norm.running_var, pose_norm.weight, pose_norm.bias)]
        eps = float(pose_norm.eps)
    ds = _f32c(det_scores) if det_scores is not None else None
    h1, h2 = w1.shape[0], w2t.shape[1]
    assert w1.shape == (h1, 17) and w2t.shape == (h1, h2) and w3.numel() == h2 and b3.numel() == 1
    assert reg is None or reg.shape == (n, h1)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_score_stage(
            ctx.ptr, _ptr(rows), _ptr(dims), _ptr(ls), float(cov_correction_sd), int(bool(distance_z_depth)),
            int(bool(use_calib)), _ptr(nm[0]), _ptr(nm[1]), _ptr(nm[2]), _ptr(nm[3]), eps, _ptr(reg), _ptr(w1), _ptr(b1),
            _ptr(w2t), _ptr(b2), _ptr(w3), _ptr(b3), h1, h2, _ptr(ds), int(bool(pre_sigmoid)), _ptr(scores), _ptr(bbox),
            _ptr(cal), _ptr(logits), n, _native.ffi.cast('void*', stream)))
    return scores, bbox, cal, logits


def gather_wait(device, flags_ptr=None, n=1, value=0, peer_acks=None, ack_slot=0, ack_value=0):
    """``mrpnp_gather_wait`` on the current stream of ``device``: wait until the ``n`` completion flags at ``flags_ptr``
    reach ``value`` and / or store ``ack_value`` into slot ``ack_slot`` of every array in ``peer_acks``."""
    dev = torch.device(device)
    ctx = get_ctx(dev)
    arr = _native.ffi.NULL
    if peer_acks:
        arr = _native.ffi.new('uint32_t*[]', [_native.ffi.cast('uint32_t*', int(q)) for q in peer_acks])
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_gather_wait(
            ctx.ptr, _native.ffi.cast('const uint32_t*', int(flags_ptr)) if flags_ptr else _native.ffi.NULL, int(n),
            int(value) & 0xffffffff, arr, int(ack_slot), int(ack_value) & 0xffffffff, _native.ffi.cast('void*', stream)))


def gather_timeouts(device=None):
    """Flag / ack waits that gave up (``mrpnp_gather_timeouts``; 0 in a healthy run).  Synchronises the device."""
    return int(_native.lib().mrpnp_gather_timeouts(get_ctx(device).ptr))


def nms_bev(bbox_3d, labels=None, group_offsets=None, iou_thr=0.25, max_group=None):
    """``mrpnp_nms_bev``: class-wise rotated BEV NMS per image.  bbox_3d [N,8] (l,h,w,x,y,z,ry,score), labels [N]
    int64 or None, group_offsets [G+1] int32 tensor or python list (None = one image).  Returns keep [N] bool.
    With a device tensor of offsets pass ``max_group`` (an upper bound of the largest image) to avoid a host sync --
    required inside CUDA-graph capture."""
    dev = bbox_3d.device
    ctx = get_ctx(dev)
    n = bbox_3d.shape[0]
    keep = torch.zeros((n,), dtype=torch.uint8, device=dev)
    if n == 0:
        return keep.bool()
    if group_offsets is None:
        group_offsets = [0, n]
    if not torch.is_tensor(group_offsets):
        sizes = [b - a for a, b in zip(group_offsets[:-1], group_offsets[1:])]
        max_group = max(sizes) if sizes else 0
        group_offsets = torch.tensor(group_offsets, dtype=torch.int32, device=dev)
    else:
        if max_group is None:
            max_group = int((group_offsets[1:] - group_offsets[:-1]).max().item())
        group_offsets = group_offsets.to(device=dev, dtype=torch.int32).contiguous()
    b = _f32c(bbox_3d)
    lab = labels.to(torch.int64).contiguous() if labels is not None else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_nms_bev(
            ctx.ptr, _ptr(b), _ptr(lab, 'int64_t*'), _ptr(group_offsets, 'int32_t*'), group_offsets.numel() - 1,
            int(max_group), float(iou_thr), _ptr(keep, 'uint8_t*'), _native.ffi.cast('void*', stream)))
    return keep.bool()


def solve_dense(noc_pred, proj_logstd, rois, dims, dims_var, cam_mats, uv_range, *, noc_mean, noc_std, focal_gain,
                scaling_denominator, distance=None, distance_min=0.1, init_pose=None, z_min=0.5, std_scale=10.0,
                istd_thres=0.6, inlier_opt_only=True, cov_mode='pipeline', precision='fast', max_iterations=50,
                return_inlier_mask=True, labels=None, num_classes=0, ransac_ratio=0.0, dim_coder=None, dim_labels=None):
    """Fused head -> PnP launch -- direct wrapper of ``mrpnp_solve_dense``: the dense head's raw class-sliced
    ``noc_pred`` [N,3,H,W] and ``proj_logstd`` [N,2,H,W], the detection boxes ``rois`` [N,4|5] and the decoded
    ``dims`` [N,3] (+ ``dims_var`` [N,3] | None, ``distance`` [N] | None) go in; NOCCoder.decode, the variance
    propagation of DistanceInvarProjErrorCoder.decode_logstd and the RoI pixel grid are evaluated in the kernel
    prologue.  Returns (result [N,24], inlier_mask [N,H*W] bool | None).

    With ``labels`` [N] int64 and ``num_classes`` = C, ``noc_pred`` is instead the head's full output ``all_pred``
    [N, 5*C, H, W] (any object stride: a ``[:, half]`` view of the flip-paired [N,2,5*C,H,W] tensor works) and the
    class slice of FCNNOCDecoder.slice_pred is taken by the kernel's loads; ``proj_logstd`` is ignored.

    With ``dim_coder`` (a MultiClassNormDimCoder) and ``dim_labels`` [N] int64, ``dims`` / ``dims_var`` are the ENCODED
    regression outputs: the kernel prologue decodes them per class (multiclass_norm_dim_coder.py:28-36) and the decoded
    tensors are returned as third and fourth element: (result, inlier_mask, dims [N,3], dims_var [N,3] | None)."""
    dev = noc_pred.device
    ctx = get_ctx(dev)
    n, _, h, w = noc_pred.shape
    n_pts = h * w
    pred_stride = 0
    if num_classes:
        if noc_pred.dtype != torch.float32 or noc_pred.shape[1] != 5 * num_classes or noc_pred[0].stride() != (h * w, w, 1):
            raise ValueError('solve_dense: all_pred must be float32 [N, 5*C, H, W] with contiguous objects')
        pred_stride = noc_pred.stride(0) if n > 1 else 5 * num_classes * n_pts
        labels = labels.to(torch.int64).contiguous()
        proj_logstd = None
    result = torch.empty((n, RESULT_STRIDE), dtype=torch.float32, device=dev)
    inl_out = torch.empty((n, (n_pts + 31) // 32), dtype=torch.int32, device=dev) if return_inlier_mask else None
    dims_dec = dims_var_dec = None
    if dim_coder is not None:
        dim_labels = (labels if dim_labels is None else dim_labels).to(torch.int64).contiguous()
        dims_dec = torch.empty((n, 3), dtype=torch.float32, device=dev)
        dims_var_dec = torch.empty((n, 3), dtype=torch.float32, device=dev) if dims_var is not None else None
    if n == 0:
        mask = unpack_mask(inl_out, n_pts) if inl_out is not None else None
        return (result, mask) if dim_coder is None else (result, mask, dims_dec, dims_var_dec)
    noc = noc_pred if num_classes else _f32c(noc_pred)
    ls = _f32c(proj_logstd) if proj_logstd is not None else None
    boxes = _f32c(rois[:, -4:])
    dm = _f32c(dims)
    dv = _f32c(dims_var) if dims_var is not None else None
    dist = _f32c(distance).reshape(-1) if distance is not None else None
    if dm.shape != (n, 3) or boxes.shape != (n, 4) or (ls is not None and ls.shape != (n, 2, h, w)) or \
            (dist is not None and dist.numel() != n):
        raise ValueError('solve_dense: inconsistent shapes')
    cam, rng = _f32c(cam_mats).reshape(-1, 9), _f32c(uv_range).reshape(-1, 4)
    if cam.shape[0] not in (1, n) or rng.shape[0] not in (1, n):
        raise ValueError('cam_mats / uv_range must have batch size 1 or N')
    init = _f32c(init_pose) if init_pose is not None else None
    p = make_params(
        n, n_pts, cam_stride=9 if cam.shape[0] == n and n > 1 else 0,
        range_stride=4 if rng.shape[0] == n and n > 1 else 0, precision=_PREC[precision],
        cov_mode={'none': 0, 'pipeline': 1, 'ceres': 2}[cov_mode],
        init_mode=C['MRPNP_INIT_GIVEN'] if init is not None else C['MRPNP_INIT_LINEAR'],
        inlier_opt_only=int(bool(inlier_opt_only)), max_iterations=int(max_iterations),
        z_min=float(z_min), std_scale=float(std_scale), istd_thres=float(istd_thres),
        ransac_ratio=float(ransac_ratio or 0.0))   # threshold = ratio * (v[last row] - v[first row]), evaluated in the kernel
    dp = _native.ffi.new('mrpnp_dense_params*')
    for i in range(3):
        dp.noc_mean[i], dp.noc_std[i] = float(noc_mean[i]), float(noc_std[i])
    dp.focal_gain, dp.scaling_denominator = float(focal_gain), float(scaling_denominator)
    dp.distance_min, dp.roi_w = float(distance_min), int(w)
    dp.num_classes, dp.pred_stride = int(num_classes), int(pred_stride)
    lab = labels if num_classes else None
    if dim_coder is not None:
        means, stds = _dim_tables(dim_coder, dev)
        dp.dim_means, dp.dim_stds, dp.n_dim_classes = _ptr(means), _ptr(stds), means.shape[0]
        dp.dims_out, dp.dims_var_out = _ptr(dims_dec), _ptr(dims_var_dec)
        if lab is not None and lab.data_ptr() != dim_labels.data_ptr() and not torch.equal(lab, dim_labels):
            raise ValueError('solve_dense: channel labels and dim_labels differ')
        lab = dim_labels
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(_native.lib().mrpnp_solve_dense(
            ctx.ptr, p, dp, _ptr(noc), _ptr(ls), _ptr(boxes), _ptr(lab, 'int64_t*'), _ptr(dm),
            _ptr(dv), _ptr(dist), _ptr(cam), _ptr(rng),
            _ptr(init), _ptr(result), _ptr(inl_out, 'uint32_t*'), _native.ffi.cast('void*', stream)))
    mask = unpack_mask(inl_out, n_pts) if inl_out is not None else None
    return (result, mask) if dim_coder is None else (result, mask, dims_dec, dims_var_dec)


_DIM_TABLES = {}


def _dim_tables(dim_coder, dev):
    """Device copies of a MultiClassNormDimCoder's per-class tables ([C,3] each), cached per (values, device)."""
    key = (tuple(map(tuple, dim_coder.target_means)), tuple(map(tuple, dim_coder.target_stds)), str(dev))
    t = _DIM_TABLES.get(key)
    if t is None:
        t = (torch.tensor(dim_coder.target_means, dtype=torch.float32, device=dev).reshape(-1, 3).contiguous(),
             torch.tensor(dim_coder.target_stds, dtype=torch.float32, device=dev).reshape(-1, 3).contiguous())
        _DIM_TABLES[key] = t
    return t


def solve_host(coords_3d, coords_2d, weights, cam_mats, uv_range, init_pose=None, *, device=0, layout='planar',
               weight_mode='logstd', result=None, **kw):
    """``mrpnp_solve_host`` on CPU tensors / numpy-backed memory (pinned for full speed): the reference op's
    host-buffer calling convention (pnp_uncert_cpu.py:128-209).  Returns result [N,24] (CPU tensor)."""
    ctx = get_ctx(torch.device('cuda', device))
    n = coords_3d.shape[0]
    n_pts = coords_3d[0].numel() // 3
    if result is None:
        result = torch.empty((n, RESULT_STRIDE), dtype=torch.float32)
    cam, rng = _f32c(cam_mats).reshape(-1, 9), _f32c(uv_range).reshape(-1, 4)
    wmode = {'logstd': C['MRPNP_W_LOGSTD'], 'istd': C['MRPNP_W_ISTD'], 'full': C['MRPNP_W_FULL']}[weight_mode]
    p = make_params(
        n, n_pts, layout=C['MRPNP_LAYOUT_PLANAR'] if layout == 'planar' else C['MRPNP_LAYOUT_INTERLEAVED'],
        weight_mode=wmode, cam_stride=9 if cam.shape[0] == n and n > 1 else 0,
        range_stride=4 if rng.shape[0] == n and n > 1 else 0,
        precision=_PREC[kw.get('precision', 'fast')],
        cov_mode={'none': 0, 'pipeline': 1, 'ceres': 2}[kw.get('cov_mode', 'pipeline')],
        init_mode=C['MRPNP_INIT_GIVEN'] if init_pose is not None else C['MRPNP_INIT_LINEAR'],
        z_min=float(kw.get('z_min', 0.5)), std_scale=float(kw.get('std_scale', 10.0)),
        istd_thres=float(kw.get('istd_thres', 0.6)))
    for t in (coords_3d, coords_2d, weights):
        assert t.device.type == 'cpu' and t.dtype == torch.float32 and t.is_contiguous()
    _native.check(_native.lib().mrpnp_solve_host(
        ctx.ptr, p, _ptr(coords_3d), _ptr(coords_2d), _ptr(weights), _ptr(cam), _ptr(rng),
        _ptr(_f32c(init_pose) if init_pose is not None else None), _native.ffi.NULL, _ptr(result),
        _native.ffi.NULL))
    return result


def u2d_pnp_cpu(coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min=0.5, epnp_istd_thres=1.0,
                epnp_ransac_thres=None, inlier_opt_only=False, with_pose_cov=True, device=0):
    """Drop-in for the module-level export ``monorun.ops.u2d_pnp_cpu`` (pnp_uncert_cpu.py:128-209): numpy fp32 arrays
    in, numpy out -- ``(ret_val (N,) bool, yaw (N,1), t_vec (N,3), pose_cov (N,4,4) | None, tr_radius (N,1),
    inlier_mask (N,P) bool)`` -- solved on CUDA device ``device`` (there is no CPU implementation behind this name;
    the arrays are copied to the device and back, like the reference copies tensors to the host).  pose_cov is Ceres'
    own covariance (``with_pose_cov``; pnp_uncert_cpu.cpp:279-291), as in the reference's native call."""
    import numpy as np
    n, p = coords_2d.shape[0], coords_2d.shape[1]
    if n == 0:   # pnp_uncert_cpu.py:201-207
        return (np.zeros((0,), bool), np.zeros((0, 1), np.float32), np.zeros((0, 3), np.float32),
                np.zeros((0, 4, 4), np.float32) if with_pose_cov else None, np.zeros((0, 1), np.float32), np.zeros((0, p), bool))
    dev = torch.device('cuda', device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)
    ur, vr = t(np.broadcast_to(u_range, (max(len(u_range), len(v_range)), 2))), t(np.broadcast_to(v_range, (max(len(u_range), len(v_range)), 2)))
    thr = t(epnp_ransac_thres) if epnp_ransac_thres is not None and inlier_opt_only else None
    result, inlier_mask, _ = solve_batched(
        t(coords_3d), t(coords_2d), t(coords_2d_istd), t(cam_mats), torch.cat([ur, vr], 1), layout='interleaved',
        weight_mode='istd', z_min=z_min, istd_thres=epnp_istd_thres, inlier_opt_only=inlier_opt_only,
        cov_mode='ceres' if with_pose_cov else 'none', ransac_thres=thr)
    r = result.cpu().numpy()
    return (r[:, 20] > 0.5, r[:, 0:1].copy(), r[:, 1:4].copy(), r[:, 4:20].reshape(n, 4, 4).copy() if with_pose_cov else None,
            r[:, 23:24].copy(), inlier_mask.cpu().numpy())


def _unpack(result, inlier_mask):
    ret_val = result[:, 20] > 0.5
    r_vec = result[:, 0:1].clone()
    t_vec = result[:, 1:4].clone()
    pose_cov = result[:, 4:20].reshape(-1, 4, 4).clone()
    return ret_val, r_vec, t_vec, pose_cov, inlier_mask


def pnp_uncert(coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min=0.5, epnp_istd_thres=1.0,
               epnp_ransac_thres=None, inlier_opt_only=False, forward_exact_hessian=False, use_6dof=False,
               init_pose=None, precision='fast'):
    """Drop-in for monorun/ops/least_squares/pnp_uncert.py:7-87.

    Args (as in the reference):
        coords_2d (Tensor): (Nbatch, Npoint, 2);  coords_2d_istd (Tensor): (Nbatch, Npoint, 2)
        coords_3d (Tensor): (Nbatch, Npoint, 3);  cam_mats (Tensor): (Nbatch, 3, 3) or (1, 3, 3)
        u_range, v_range (Tensor): (Nbatch, 2) or (1, 2);  z_min, epnp_istd_thres (float)
        epnp_ransac_thres (None | Tensor): (Nbatch,) reprojection threshold in pixels.  The reference hands it to
            cv2.solvePnPRansac, whose consensus set narrows the inliers LM, the covariance and the returned mask see
            (pnp_uncert_cpu.py:34-51).  Here the start pose (on-device linear initialiser, or init_pose) is the model
            of a deterministic consensus pass on the device with the same rule (dropped only if > 4 points survive);
            OpenCV's random sampling itself is not reproduced.
        forward_exact_hessian: True replaces the Gauss-Newton covariance by the inverse of the second-order
            Hessian (hessian.py:5-64; one more launch, ``mrpnp_exact_hessian``).  use_6dof: accepted and unused,
            exactly as in the reference.
        init_pose (Tensor | None): extension -- (Nbatch, 4) [yaw, t] to start LM from (e.g. an EPnP result).
    Returns:
        ret_val (Nbatch,) bool, r_vec (Nbatch, 1), t_vec (Nbatch, 3), pose_cov (Nbatch, 4, 4),
        inlier_mask (Nbatch, Npoint) bool -- all on the input device.
    """
    with torch.no_grad():
        n = coords_2d.shape[0]
        if n == 0:  # pnp_uncert.py:60-61, pnp_uncert_cpu.py:201-207
            return (coords_2d.new_zeros((0,), dtype=torch.bool), coords_2d.new_zeros((0, 1)),
                    coords_2d.new_zeros((0, 3)), coords_2d.new_zeros((0, 4, 4)),
                    coords_2d.new_zeros((0, coords_2d.shape[1]), dtype=torch.bool))
        uv_range = torch.cat([u_range.expand(max(u_range.shape[0], v_range.shape[0]), 2),
                              v_range.expand(max(u_range.shape[0], v_range.shape[0]), 2)], dim=1)
        result, inlier_mask, _ = solve_batched(
            coords_3d, coords_2d, coords_2d_istd, cam_mats, uv_range, init_pose=init_pose, layout='interleaved',
            weight_mode='istd', z_min=z_min, istd_thres=epnp_istd_thres, inlier_opt_only=inlier_opt_only,
            cov_mode='none' if forward_exact_hessian else 'pipeline', precision=precision,
            ransac_thres=epnp_ransac_thres if inlier_opt_only else None)
        if forward_exact_hessian:  # pnp_uncert.py:63-69, :77-85
            exact_hessian(coords_3d, coords_2d, coords_2d_istd, cam_mats, uv_range, result, inlier_mask,
                          layout='interleaved', weight_mode='istd', z_min=z_min, rows=result, return_hessian=False)
        return _unpack(result, inlier_mask)


@PNP.register_module()
class PnPUncert(torch.nn.Module):
    """Drop-in for monorun/ops/least_squares/pnp_uncert.py:90-142 (same constructor kwargs and forward)."""

    def __init__(self, z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True, coord_istd_normalize=False,
                 forward_exact_hessian=False, use_6dof=False, eps=1e-6, precision='fast'):
        super(PnPUncert, self).__init__()
        self.z_min = z_min
        self.epnp_istd_thres = epnp_istd_thres
        self.inlier_opt_only = inlier_opt_only
        self.coord_istd_normalize = coord_istd_normalize
        self.forward_exact_hessian = forward_exact_hessian
        self.use_6dof = use_6dof
        self.eps = eps
        self.precision = precision

    def forward(self, coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, epnp_ransac_thres=None,
                init_pose=None):
        if self.coord_istd_normalize:  # pnp_uncert.py:130-132
            mean = torch.mean(coords_2d_istd, dim=(1, 2), keepdim=True)
            coords_2d_istd = coords_2d_istd / mean.clamp(min=self.eps)
        return pnp_uncert(
            coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min=self.z_min,
            epnp_istd_thres=self.epnp_istd_thres, epnp_ransac_thres=epnp_ransac_thres,
            inlier_opt_only=self.inlier_opt_only, forward_exact_hessian=self.forward_exact_hessian,
            use_6dof=self.use_6dof, init_pose=init_pose, precision=self.precision)

    def forward_6dof(self, coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, init_pose=None):
        """``use_6dof=True`` extension (the reference accepts the flag and never reads it, pnp_uncert.py:11,98,122,142):
        the 4-DoF solve of :meth:`forward` supplies the start (0, yaw, 0, t) and the inlier mask, then all six pose
        parameters are refined.  Returns (ret_val (N,), r_vec (N,3) angle-axis, t_vec (N,3), pose_cov (N,6,6),
        inlier_mask (N,P))."""
        with torch.no_grad():
            ret_val, yaw, t_vec, _, inlier_mask = self.forward(coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range,
                                                               v_range, init_pose=init_pose)
            n = coords_2d.shape[0]
            if n == 0:
                return ret_val, coords_2d.new_zeros((0, 3)), t_vec, coords_2d.new_zeros((0, 6, 6)), inlier_mask
            zero = torch.zeros_like(yaw)
            init6 = torch.cat([zero, yaw, zero, t_vec], 1)
            rows = max(u_range.shape[0], v_range.shape[0])
            uv_range = torch.cat([u_range.expand(rows, 2), v_range.expand(rows, 2)], dim=1)
            istd = coords_2d_istd
            if self.coord_istd_normalize:
                istd = istd / torch.mean(istd, dim=(1, 2), keepdim=True).clamp(min=self.eps)
            res = solve_6dof_batched(coords_3d, coords_2d, istd, cam_mats, uv_range, init6,
                                     inlier_mask if self.inlier_opt_only else None, layout='interleaved',
                                     weight_mode='istd', z_min=self.z_min, precision=self.precision)
            ok = ret_val & (res[:, 42] > 0.5)
            return (ok, res[:, 0:3].float(), res[:, 3:6].float(), res[:, 6:42].reshape(n, 6, 6).float(), inlier_mask)

    def forward_dense(self, coords_2d, coords_2d_logstd, coords_3d, cam_mats, uv_range, std_scale, init_pose=None,
                      epnp_ransac_thres=None):
        """Head-level entry used by UncertPropPnPOptimizer: NCHW tensors and log-std straight into the kernel
        (fuses uncert_prop_pnp_optimizer.py:73 and the three permute copies of :82-84)."""
        with torch.no_grad():
            if self.coord_istd_normalize:
                raise NotImplementedError('coord_istd_normalize with the dense entry')
            result, inlier_mask, _ = solve_batched(
                coords_3d, coords_2d, coords_2d_logstd, cam_mats, uv_range, init_pose=init_pose, layout='planar',
                weight_mode='logstd', z_min=self.z_min, std_scale=std_scale, istd_thres=self.epnp_istd_thres,
                inlier_opt_only=self.inlier_opt_only,
                cov_mode='none' if self.forward_exact_hessian else 'pipeline', precision=self.precision,
                ransac_thres=epnp_ransac_thres if self.inlier_opt_only else None)
            if self.forward_exact_hessian:
                exact_hessian(coords_3d, coords_2d, coords_2d_logstd, cam_mats, uv_range, result, inlier_mask,
                              layout='planar', weight_mode='logstd', z_min=self.z_min, std_scale=std_scale,
                              rows=result, return_hessian=False)
            return _unpack(result, inlier_mask)

    def forward_fused(self, noc_pred, proj_logstd, rois, dims, dims_var, cam_mats, uv_range, std_scale, coord_coder,
                      proj_error_coder, distance=None, init_pose=None, labels=None, num_classes=0, ransac_ratio=0.0,
                      dim_coder=None, dim_labels=None):
        """Fused head -> PnP entry (``mrpnp_solve_dense``): takes what FCNNOCDecoder returns plus the decoded
        dimensions and the boxes; ``coord_coder`` (NOCCoder) and ``proj_error_coder``
        (DistanceInvarProjErrorCoder) only supply their constants.  With ``labels`` / ``num_classes`` the first
        argument is the head's unsliced ``all_pred``; with ``dim_coder`` / ``dim_labels`` the dimensions come in encoded
        and the decoded (dims, dims_var) are appended to the five outputs (see :func:`solve_dense`)."""
        with torch.no_grad():
            if self.coord_istd_normalize:
                raise NotImplementedError('coord_istd_normalize with the fused entry')
            if self.forward_exact_hessian:
                raise NotImplementedError('forward_exact_hessian with the fused entry (the decoded tensors it needs '
                                          'are never materialised); use forward_dense')
            result, inlier_mask, *decoded = solve_dense(
                noc_pred, proj_logstd, rois, dims, dims_var, cam_mats, uv_range, dim_coder=dim_coder, dim_labels=dim_labels,
                noc_mean=coord_coder.target_means, noc_std=coord_coder.target_stds,
                focal_gain=proj_error_coder.ref_focal_y * proj_error_coder.epistemic_std_gain,
                scaling_denominator=proj_error_coder.scaling_denomitor, distance=distance,
                distance_min=proj_error_coder.distance_min, init_pose=init_pose, z_min=self.z_min,
                std_scale=std_scale, istd_thres=self.epnp_istd_thres, inlier_opt_only=self.inlier_opt_only,
                cov_mode='pipeline', precision=self.precision, labels=labels, num_classes=num_classes,
                ransac_ratio=ransac_ratio if self.inlier_opt_only else 0.0)
            return (*_unpack(result, inlier_mask), *decoded)
