"""`utils` as the reference's train.py imports it (train.py:12-15) -> the B200 helpers."""
from ideas_b200.utils import (accumulate, d_logistic_loss, d_r1_loss, data_sampler, g_nonsaturating_loss,  # noqa: F401
                              message_to_tensor, patchify_image, requires_grad, sample_data, tensor_to_message,
                              time_change)
