"""`from net.prompt_xrestormer import PromptXRestormer` -- the reference's import line -- resolves to the B200 implementation."""
from promptir_b200.net.prompt_xrestormer import PromptXRestormer  # noqa: F401

__all__ = ["PromptXRestormer"]
