"""B200-native EDM sampling / denoising hot path with the reference's Hydra-facing class API.

Drop-in `_target_`s (reference path -> this package):
    src.models.components.diffusion.EluDiffusion          -> audiodiffuser_b200.components.diffusion.EluDiffusion
    src.models.components.sampler_edm.EDMSampler          -> audiodiffuser_b200.components.sampler_edm.EDMSampler
    src.models.components.sampler_edm.EDMAlphaSampler     -> audiodiffuser_b200.components.sampler_edm.EDMAlphaSampler
    src.models.components.scheduler.KarrasSchedule        -> audiodiffuser_b200.components.scheduler.KarrasSchedule
    src.models.components.distribution.LogNormalDistribution -> audiodiffuser_b200.components.distribution.LogNormalDistribution
    src.models.backbones.wavenet.WaveNetNoise             -> audiodiffuser_b200.backbones.wavenet.WaveNetNoise
    src.models.backbones.unet1d.UNet1dBase                -> audiodiffuser_b200.backbones.unet1d.UNet1dBase
    src.models.components.sampler_edm.DPM2MSampler        -> audiodiffuser_b200.components.sampler_edm.DPM2MSampler
    src.models.components.stochastic_sampler_edm.ADPM2Sampler -> audiodiffuser_b200.components.sampler_edm.ADPM2Sampler
    src.models.phema.{PowerFunctionEMA,TraditionalEMA}    -> audiodiffuser_b200.ema.{PowerFunctionEMA,TraditionalEMA}
"""
from .components.diffusion import EluDiffusion, Diffusion              # noqa: F401
from .components.sampler_edm import EDMSampler, EDMAlphaSampler, DPM2MSampler, ADPM2Sampler   # noqa: F401
from .components.scheduler import KarrasSchedule                        # noqa: F401
from .components.distribution import LogNormalDistribution              # noqa: F401
from .backbones.wavenet import WaveNetNoise, EDMDenoiser                # noqa: F401
from .backbones.unet1d import UNet1dBase                                # noqa: F401

__version__ = "0.1.0"
