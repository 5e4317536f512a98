"""B200-native photometric self-supervision loss (the hot path of
MariBax/self-supervised-depth-estimation: trainer.py:465-622 over layers.py:16-25,139-248).

Hand-written sm_100a CUDA kernels behind the C-ABI library ``libpml.so`` (include/pml.h),
wrapped in ``torch.autograd.Function``s with the reference's constructor / forward signatures.
There is no CPU fallback: every op raises if the CUDA library is missing or a tensor is not on
a CUDA device.  Import as ``import ssde_b200`` (alias module at the repo root).
"""
from . import synthetic  # noqa: F401
from . import _cabi  # noqa: F401
from . import functional  # noqa: F401
from . import layers  # noqa: F401
from . import trainer_hooks  # noqa: F401
from . import hostio  # noqa: F401
from .layers import (BackprojectDepth, Project3D, SSIM, disp_to_depth,  # noqa: F401
                     get_smooth_loss, transformation_from_parameters)
from .trainer_hooks import (generate_images_pred, compute_reprojection_loss,  # noqa: F401
                            compute_losses, install)

__all__ = ["synthetic", "functional", "layers", "trainer_hooks", "BackprojectDepth", "Project3D",
           "SSIM", "disp_to_depth", "get_smooth_loss", "transformation_from_parameters",
           "generate_images_pred", "compute_reprojection_loss", "compute_losses", "install"]
