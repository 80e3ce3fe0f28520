"""Drop-in counterparts of the reference's ``utils`` package (buffer / normalization / runner / model / env), backed
by the device env and the C-ABI kernels.  Importing ``gymrl_b200.utils.install()`` registers these modules under the
reference's import names (``utils.buffer`` ...) so a legacy script's ``from utils.runner import *`` resolves here."""
import sys


def install():
    """Make `import utils.buffer / utils.runner / utils.model / utils.normalization` resolve to this package."""
    from . import buffer, model, normalization, runner  # noqa: F401
    pkg = sys.modules[__name__]
    sys.modules["utils"] = pkg
    for name in ("buffer", "model", "normalization", "runner"):
        sys.modules["utils." + name] = sys.modules[__name__ + "." + name]
    return pkg
