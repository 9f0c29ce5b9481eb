"""Drop-in mirror of the reference's ``vq`` package for the BigCodec hot path
(vq/__init__.py:1-2 exports BigCodecEncoder / BigCodecDecoder)."""
from .codec_encoder import BigCodecEncoder
from .codec_decoder import BigCodecDecoder
from .residual_vq import ResidualVQ
from .factorized_vector_quantize import FactorizedVectorQuantize
from .finite_scalar_quantization import FSQ
from .module import (CausalConv1d, CausalConvTranspose1d, DecoderBlock, EncoderBlock, ResidualUnit, ResLSTM,
                     WNConv1d, WNConvTranspose1d, get_precision, precision_scope, set_precision)
from .activations import SnakeBeta
from .alias_free_torch import Activation1d
