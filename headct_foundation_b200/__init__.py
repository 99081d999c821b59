"""headct_foundation_b200 -- B200-native (sm_100a) implementation of the HeadCT-Foundation 3-D ViT hot path.

The classes below keep the reference's `src/models` surface (constructor kwargs, methods, state_dict layout);
their compute runs in hand-written CUDA behind the C ABI in include/hct_b200.h.
"""
from .models.mae import MaskedAutoencoderViT
from .models.vit import ViT
from .models.attentionblock import AttentionBlock, SelfAttention, MLPBlock
from .models.dino_head import DINOHead
from .models.classifier import LinearClassifier, AttentionClassifier
from .models.layers import RMSNorm
from .models.attentionblock import LoraLinear
from .utils.patch_embedding import PatchEmbeddingBlock
from .utils.pos_embed import build_sincos_position_embedding
from .utils.misc import MultiCropWrapper, _update_momentum_encoder, update_momentum_encoder, set_requires_grad_false
from .utils.checkpoint import save_checkpoint, load_model, load_optimizer, interpolate_pos_embed, strip_wrapper_prefixes
from .utils.graphs import GraphedForward, GraphedTrainStep
from .utils.metrics import DownstreamMetrics, multiclass_accuracy, multiclass_auroc
from .losses.losses import DINOLoss
from .functional import set_precision, get_precision, precision
from .data.transforms import MultipleWindowScaleStack, MAE3DTrainAugment, ViTTrainAugment, DataAugmentationDINO3D

__all__ = ["MaskedAutoencoderViT", "ViT", "AttentionBlock", "SelfAttention", "MLPBlock", "DINOHead",
           "LinearClassifier", "AttentionClassifier", "RMSNorm", "LoraLinear", "PatchEmbeddingBlock", "build_sincos_position_embedding", "MultiCropWrapper",
           "update_momentum_encoder", "set_requires_grad_false", "GraphedForward", "GraphedTrainStep", "DownstreamMetrics", "multiclass_accuracy", "multiclass_auroc", "DINOLoss", "MultipleWindowScaleStack", "save_checkpoint", "load_model",
           "load_optimizer", "interpolate_pos_embed", "strip_wrapper_prefixes", "MAE3DTrainAugment", "ViTTrainAugment",
           "DataAugmentationDINO3D", "set_precision", "get_precision", "precision"]
