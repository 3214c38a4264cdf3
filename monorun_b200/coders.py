"""Elementwise decoders between the dense head and PnP, same classes/registries as the
reference so ``dict(type='NOCCoder', ...)`` blocks from configs/kitti_*.py build unchanged.

* NOCCoder.decode                        <- monorun/core/bbox_3d/coord_coder/noc_coder.py:50-73
* DistanceInvarProjErrorCoder            <- monorun/core/bbox_3d/proj_error_coder/distance_invar_proj_error_coder.py:8-63
* MultiClassNormDimCoder.decode          <- monorun/core/bbox_3d/dim_coder/multiclass_norm_dim_coder.py:28-36
* Vec2DRotationCoder                     <- monorun/core/bbox_3d/rotation_coder/vec2d_rotation_coder.py
* coords_2d_from_rois                    <- roi_align(coord_2d, rois, (28,28), 1.0, 0, 'avg', True),
                                            monorun/models/roi_heads/monorun_roi_head.py:521-523
"""
import torch

from .registry import COORD_CODERS, DIM_CODERS, PROJ_ERROR_CODERS, ROTATION_CODERS


@COORD_CODERS.register_module()
class NOCCoder(object):

    def __init__(self, target_means=(-0.1, -0.5, 0.0), target_stds=(0.35, 0.23, 0.34), eps=1e-5):
        self.target_means = target_means
        self.target_stds = target_stds
        self.eps = eps

    def decode(self, part, part_var, dimensions, dimensions_var, flip):
        """noc_coder.py:50-73: coords_3d = (noc*std+mean)*dims and first-order variance propagation."""
        dimensions = dimensions[..., None, None]
        if dimensions_var is not None:
            dimensions_var = dimensions_var[..., None, None]
        target_means = part.new_tensor(self.target_means)[:, None, None]
        target_stds = part.new_tensor(self.target_stds)[:, None, None]
        part_norm = part * target_stds + target_means
        coords_3d = part_norm * dimensions
        if part_var is not None:
            part_norm_var = part_var * target_stds.square()
            coords_3d_var = part_norm_var * dimensions.square()
            if dimensions_var is not None:
                coords_3d_var = coords_3d_var + dimensions_var * part_norm.square() + part_norm_var * dimensions_var
        elif dimensions_var is not None:
            coords_3d_var = dimensions_var * part_norm.square()
        else:
            coords_3d_var = None
        return coords_3d, coords_3d_var


@PROJ_ERROR_CODERS.register_module()
class DistanceInvarProjErrorCoder(object):

    def __init__(self, ref_length=1.6, ref_focal_y=722, target_std=0.25, distance_min=0.1, epistemic_std_gain=1.0):
        self.scaling_denomitor = ref_length * ref_focal_y * target_std  # (sic) reference attribute name
        self.ref_focal_y = ref_focal_y
        self.distance_min = distance_min
        self.epistemic_std_gain = epistemic_std_gain

    def decode_logstd(self, proj_logstd, coords_3d_var, distance):
        """distance_invar_proj_error_coder.py:39-60."""
        distance_ = distance[..., None, None].clamp(min=self.distance_min) if distance is not None \
            else proj_logstd.new_tensor([self.scaling_denomitor])
        if coords_3d_var is not None:
            coords_2d_var = torch.stack(
                [0.5 * (coords_3d_var[:, 0] + coords_3d_var[:, 2]), coords_3d_var[:, 1]], dim=1)
            coords_2d_var = (coords_2d_var * (self.ref_focal_y * self.epistemic_std_gain) ** 2
                             + (2 * proj_logstd).exp() * self.scaling_denomitor ** 2) / distance_.square()
            return 0.5 * torch.log(coords_2d_var)
        return proj_logstd + torch.log(self.scaling_denomitor / distance_)

    def cov_correction(self, cov, distance):
        """distance_invar_proj_error_coder.py:62-63."""
        return cov * (self.scaling_denomitor / distance).square().view(-1, 1, 1)


@DIM_CODERS.register_module()
class MultiClassNormDimCoder(object):

    def __init__(self,
                 target_means=[(3.89, 1.53, 1.62), (0.82, 1.78, 0.63), (1.77, 1.72, 0.57)],
                 target_stds=[(0.44, 0.14, 0.11), (0.25, 0.13, 0.12), (0.15, 0.10, 0.14)]):
        assert len(target_means) == len(target_stds)
        self.target_means = target_means
        self.target_stds = target_stds

    def decode(self, dim, dim_var, labels):
        """multiclass_norm_dim_coder.py:28-36."""
        target_means = dim.new_tensor(self.target_means)[labels]
        target_stds = dim.new_tensor(self.target_stds)[labels]
        dimensions = dim * target_stds + target_means
        dimensions_var = dim_var * target_stds.square() if dim_var is not None else None
        return dimensions, dimensions_var


@ROTATION_CODERS.register_module()
class Vec2DRotationCoder(object):
    """vec2d_rotation_coder.py:6-23: yaw -> (cos, sin); only needed so the pose_head config block builds."""

    @staticmethod
    def encode(angles):
        if len(angles.shape) == 1:
            angles = angles.unsqueeze(-1)
        return torch.cat((torch.cos(angles), torch.sin(angles)), dim=-1)

    @staticmethod
    def decode(vecs):
        raise NotImplementedError


def coords_2d_from_rois(rois, out_size=28):
    """Analytic replacement of ``roi_align(coord_2d, rois, (28,28), 1.0, 0, 'avg', True)``
    (monorun_roi_head.py:521-523): RoIAlign(aligned=True) of the pixel-index grid (u = column, v = row;
    datasets/pipelines/loading.py:67-78) returns the bin centres

        u[j] = x1 - 0.5 + (j + 0.5) (x2 - x1) / S,      v[i] = y1 - 0.5 + (i + 0.5) (y2 - y1) / S

    exactly for boxes inside the image (bilinear interpolation is exact on a linear field).  Boxes hanging more
    than one pixel over the top/left border deviate in mmcv (out-of-map samples are zeroed); detection boxes are
    clipped to the image upstream, so this only matters within 0.5 px of that border.

    rois: (N, 5) [batch_idx, x1, y1, x2, y2] or (N, 4).  Returns (N, 2, S, S) float32.
    """
    b = rois[:, -4:].float()
    k = (torch.arange(out_size, device=rois.device, dtype=torch.float32) + 0.5) * (1.0 / out_size)
    u = b[:, 0:1] - 0.5 + k[None, :] * (b[:, 2:3] - b[:, 0:1])
    v = b[:, 1:2] - 0.5 + k[None, :] * (b[:, 3:4] - b[:, 1:2])
    n = b.shape[0]
    return torch.stack([u[:, None, :].expand(n, out_size, out_size), v[:, :, None].expand(n, out_size, out_size)], dim=1)
