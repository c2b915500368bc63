"""hierarchicalprobabilistic3dhuman_b200 -- B200-native (sm_100a) implementation of the
probabilistic-pose inference hot path of akashsengupta1997/HierarchicalProbabilistic3DHuman:
ResNet-18 proxy-rep encoder -> hierarchical matrix-Fisher head -> matrix-Fisher sampler -> SMPL.

Public surface mirrors the reference's Python interface for that path:
  PoseMFShapeGaussianNet(smpl_parents, config).forward(input, input_feats=None)
  SMPL(model_path, batch_size, gender, num_betas).forward(betas, body_pose, global_orient, pose2rot)
  pose_matrix_fisher_sampling_torch(pose_U, pose_S, pose_V, num_samples, b, oversampling_ratio, sample_on_cpu)
  compute_vertex_uncertainties_by_poseMF_shapeGaussian_sampling(...)
  rot6d_to_rotmat(x)
  CannyEdgeDetector(...).forward(img), convert_2Djoints_to_gaussian_heatmaps_torch(joints2D, img_wh, std)
All arithmetic runs in hand-written CUDA behind the C ABI of include/hp3d.h (libhp3d.so)."""
from .pose_net import PoseMFShapeGaussianNet
from .smpl import SMPL, SMPLOutput
from .sampling import (pose_matrix_fisher_sampling_torch, compute_vertex_uncertainties_by_poseMF_shapeGaussian_sampling,
                       sample_meshes_batched, vertex_uncertainty, rank_samples_by_joints2d,
                       joints2D_error_sorted_verts_sampling, check_sampler_status, SamplerShortfall)
from .rigid import rot6d_to_rotmat
from .pipeline import HotPathPipeline
from .proxy import (CannyEdgeDetector, convert_2Djoints_to_gaussian_heatmaps_torch, proxy_representation,
                    joints2d_heatmap_argmax)

__all__ = ["PoseMFShapeGaussianNet", "SMPL", "SMPLOutput", "pose_matrix_fisher_sampling_torch",
           "compute_vertex_uncertainties_by_poseMF_shapeGaussian_sampling", "sample_meshes_batched",
           "vertex_uncertainty", "rot6d_to_rotmat", "HotPathPipeline", "rank_samples_by_joints2d",
           "joints2D_error_sorted_verts_sampling", "CannyEdgeDetector", "convert_2Djoints_to_gaussian_heatmaps_torch",
           "proxy_representation", "joints2d_heatmap_argmax", "check_sampler_status", "SamplerShortfall"]
