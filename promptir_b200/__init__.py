"""promptir_b200 -- B200 (sm_100a) native execution of the PromptIR restoration forward.

    from promptir_b200 import PromptIR          # drop-in for `from net.model import PromptIR`

Package layout: csrc/ (CUDA kernels + C ABI, built into libpromptir_b200.so), _lib.py/ops.py (ctypes binding),
packing.py (derived weight layouts), engine.py (buffer plan + launch program), net/model.py (the nn.Module),
tiling.py (batched tile_eval), train_engine.py (forward + hand-written backward programs), ddp.py (flat-gradient all-reduce),
net/prompt_xrestormer.py + xengine.py (the PromptXRestormer variant).
"""
from .net.model import PromptIR  # noqa: F401
from .net.prompt_xrestormer import PromptXRestormer  # noqa: F401

__all__ = ["PromptIR", "PromptXRestormer"]
