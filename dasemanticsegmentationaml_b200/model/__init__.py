from .stdcnet import STDCNet813, ConvX, CatBottleneck, AddBottleneck  # noqa: F401
from .model_stages import (BiSeNet, ContextPath, AttentionRefinementModule, FeatureFusionModule,  # noqa: F401
                           BiSeNetOutput, ConvBNReLU)
from .discriminator import (FCDiscriminator, DepthWiseSepFCDiscriminator,  # noqa: F401
                            DepthWiseSepBNFCDiscriminator)
