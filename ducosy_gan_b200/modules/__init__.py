from .model import (CBAM, ChannelAttention, Discriminator, Generator, ResidualBlock, ResidualBlockWithCBAM,
                    SpatialAttention, weights_init_normal)

__all__ = ["Generator", "Discriminator", "weights_init_normal", "ChannelAttention", "SpatialAttention", "CBAM",
           "ResidualBlock", "ResidualBlockWithCBAM"]
