from .custom_clip import ClipTestTimeTuning, LoRA_AB, get_coop  # noqa: F401
