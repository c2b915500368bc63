"""Drop-in for the matrix-Fisher normalising constant of the reference's pose NLL:

  LogMFNormConstant.apply(S)     -- reference losses/matrix_fisher_loss.py:134-192 (autograd Function: forward log c(S),
                                    backward d log c / d S)
  matrix_fisher_expected_rotation(U, S, V)  -- E[R] = U_p diag(d log c / d s) V_p^T of the distribution the sampler draws from

One kernel launch computes log c and its gradient (csrc/mf_norm.cu); the backward pass only scales the saved gradient.
STATUS: arithmetic (csrc/mf_norm_math.h) verified on the host against the reference-pinned oracle; the CUDA kernel has not
yet run on hardware."""
import torch

from . import _lib


def _log_norm_and_grad(S):
    _lib.require_cuda(S, "S")
    s = S.detach().to(torch.float32).contiguous().view(-1, 3)
    n = s.shape[0]
    log_c = torch.empty(n, device=s.device, dtype=torch.float32)
    grad = torch.empty(n, 3, device=s.device, dtype=torch.float32)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().hp3d_mf_log_norm_constant(s.data_ptr(), n, log_c.data_ptr(), grad.data_ptr(), _lib.stream_ptr()),
                   "hp3d_mf_log_norm_constant")
    return log_c, grad


class LogMFNormConstant(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S):
        """S (B,3) proper singular values ordered big to small -> log c(S) (B,)."""
        log_c, grad = _log_norm_and_grad(S)
        ctx.save_for_backward(grad)
        return log_c

    @staticmethod
    def backward(ctx, grad_log_c):
        (grad,) = ctx.saved_tensors
        return (grad * grad_log_c.view(-1, 1)).view(-1, 3)


def matrix_fisher_expected_rotation(pose_U, pose_S, pose_V):
    """(...,3,3), (...,3), (...,3,3) improper SVD factors as the head returns them -> E[R] (...,3,3)."""
    U = pose_U.reshape(-1, 3, 3).clone()
    V = pose_V.reshape(-1, 3, 3).clone()
    S = pose_S.reshape(-1, 3).clone()
    du, dv = torch.linalg.det(U), torch.linalg.det(V)
    U[:, :, 2] *= du[:, None]
    V[:, :, 2] *= dv[:, None]
    S[:, 2] *= du * dv
    _, d = _log_norm_and_grad(S)
    return (U * d[:, None, :]) @ V.transpose(1, 2)
