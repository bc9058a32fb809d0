"""`utils` as the reference's train.py imports it (train.py:12-15) -> the B200 helpers."""
from ideas_b200 import utils as _U
from ideas_b200.utils import (accumulate, d_logistic_loss, d_r1_loss, data_sampler, g_nonsaturating_loss,  # noqa: F401
                              message_to_tensor, patchify_image, requires_grad, sample_data, time_change)


def tensor_to_message(secret_tensor, sigma):
    """Reference placement (utils.py:86-97): the message is built with a default-device ``torch.zeros``, i.e. it
    comes back on the CPU whatever the device of the secret tensor -- train.py:285 subtracts it from the CPU
    message M.  The decode itself is the integer CUDA kernel; only the unpacked bits cross to the host."""
    return _U.tensor_to_message(secret_tensor, sigma).cpu()
