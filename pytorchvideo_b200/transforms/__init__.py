from .transforms import (ApplyTransformToKey, CenterCropVideo, ConvertUint8ToFloat, Div255,  # noqa: F401
                         FusedClipTransform, Normalize, RandomCropVideo, RandomShortSideScale, ShortSideScale,
                         UniformCropVideo, UniformTemporalSubsample, create_video_transform, SlowFastPackPathway,
                         RemoveKey)
from . import functional  # noqa: F401
