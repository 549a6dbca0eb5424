"""Model helpers with the reference's names (/root/reference/compressai/models/utils.py:46-132)."""
import torch
import torch.nn as nn


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


def find_named_buffer(module, query):
    return next((b for n, b in module.named_buffers() if n == query), None)


def _update_registered_buffer(module, buffer_name, state_dict_key, state_dict, policy="resize_if_empty", dtype=torch.int):
    new_size = state_dict[state_dict_key].size()
    registered = find_named_buffer(module, buffer_name)
    if policy in ("resize_if_empty", "resize"):
        if registered is None:
            raise RuntimeError(f'buffer "{buffer_name}" was not registered')
        if policy == "resize" or registered.numel() == 0:
            registered.resize_(new_size)
    elif policy == "register":
        if registered is not None:
            raise RuntimeError(f'buffer "{buffer_name}" was already registered')
        module.register_buffer(buffer_name, torch.empty(new_size, dtype=dtype).fill_(0))
    else:
        raise ValueError(f'Invalid policy "{policy}"')


def update_registered_buffers(module, module_name, buffer_names, state_dict, policy="resize_if_empty", dtype=torch.int):
    """Resize the (empty) CDF buffers of `module` to the sizes found in a checkpoint before loading it."""
    valid = [n for n, _ in module.named_buffers()]
    for name in buffer_names:
        if name not in valid:
            raise ValueError(f'Invalid buffer name "{name}"')
    for name in buffer_names:
        key = f"{module_name}.{name}"
        if key in state_dict:
            _update_registered_buffer(module, name, key, state_dict, policy, dtype)
