"""gymtorch: tensors handed out by the fake gym already are torch tensors."""


def wrap_tensor(t):
    return t


def unwrap_tensor(t):
    return t
