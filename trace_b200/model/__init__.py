from .trace_mistral import TraceMistralConfig, TraceMistralForCausalLM  # noqa: F401
