from .gpt2 import GPT2InferenceEngine  # noqa: F401
