"""`models` as the reference's train.py imports it (train.py:17) -> the B200 networks."""
from ideas_b200.models import *  # noqa: F401,F403
from ideas_b200.models import init_model  # noqa: F401
