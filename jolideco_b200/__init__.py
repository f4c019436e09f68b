"""jolideco_b200 — B200-native (sm_100a) implementation of Jolideco's MAP deconvolution hot path,
drop-in behind the reference's Python API.  All arithmetic runs in `libjolideco_b200.so`
(hand-written CUDA, C ABI in include/jolideco_b200.h); there is no CPU or PyTorch fallback."""
from ._lib import JolidecoB200Error  # noqa: F401
from .batch import BatchedRuns, run_many  # noqa: F401
from .core import MAPDeconvolver, MAPDeconvolverResult  # noqa: F401
from .loss import PoissonLoss, PriorLoss, TotalLoss  # noqa: F401
from .models import (  # noqa: F401
    FluxComponents,
    NPredCalibration,
    NPredCalibrations,
    NPredModel,
    NPredModels,
    SpatialFluxComponent,
)
from .norms import (  # noqa: F401
    NORMS_REGISTRY,
    ASinhImageNorm,
    ATanImageNorm,
    FixedMaxImageNorm,
    IdentityImageNorm,
    ImageNorm,
    InverseCDFImageNorm,
    LogImageNorm,
    MaxImageNorm,
    PowerImageNorm,
    SigmoidImageNorm,
)
from .priors import (  # noqa: F401
    GaussianMixtureModel,
    GaussianMixtureModelMeta,
    GMMPatchPrior,
    Prior,
    Priors,
    UniformPrior,
    set_default_backend,
)

__version__ = "0.1.0"
