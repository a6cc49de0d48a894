"""Constants of the TRACE surface (mirrors trace/constants.py:6-57; values are part of the observable
behaviour: sentinel ids appear in tokenised prompts, NUM_FRAMES/MAX_FRAMES bound the frame sampler)."""
NUM_FRAMES = 8
MAX_FRAMES = 128
NUM_FRAMES_PER_SECOND = 1

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_VIDEO_TOKEN = "<video>"

MMODAL_TOKEN_INDEX = {"IMAGE": -200, "VIDEO": -201, "AUDIO": -202, "TIME": -203, "SCORE": -204, "SYNC": -205}
MMODAL_INDEX_TOKEN = {v: k for k, v in MMODAL_TOKEN_INDEX.items()}
DEFAULT_MMODAL_TOKEN = {"IMAGE": "<image>", "VIDEO": "<video>", "AUDIO": "<audio>", "TIME": "<time>",
                        "SCORE": "<score>", "SYNC": "<sync>"}
