"""Checkpoint key fix-ups (reference: /root/reference/compressai/zoo/pretrained.py:19-50)."""


_EB_LISTS = {"_biases": "_bias", "_matrices": "_matrix", "_factors": "_factor"}


def rename_key(key):
    """Strip DataParallel prefixes and map old EntropyBottleneck ParameterList names to the flat ones."""
    if key.startswith("module."):
        key = key[7:]
    if key.startswith("h_s."):
        return None
    for plural, singular in _EB_LISTS.items():
        tag = f"entropy_bottleneck.{plural}."
        if key.startswith(tag):
            return f"entropy_bottleneck.{singular}{key[len(tag):]}"
    return key


def load_pretrained(state_dict):
    out = {}
    for k, v in state_dict.items():
        nk = rename_key(k)
        if nk is not None:
            out[nk] = v
    return out
