"""Import shim so the reference's scripts (`from net.model import PromptIR`, train.py:10, test.py:14, demo.py:8)
pick up the B200 implementation unchanged when this repository is first on sys.path."""
from promptir_b200.net.model import (Attention, BiasFree_LayerNorm, Downsample, FeedForward, LayerNorm,  # noqa: F401
                                     OverlapPatchEmbed, PromptGenBlock, PromptIR, TransformerBlock, Upsample,
                                     WithBias_LayerNorm, resblock)
