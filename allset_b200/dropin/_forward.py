"""Shared loader for the drop-in `layers` / `models` modules: load the reference's own file under a private name and
re-export everything it defines, so `from layers import *` / `from models import *` in the reference's train.py
(reference src/train.py:21-23) keep resolving every baseline model, while the hot-path classes are replaced by the
B200-native ones.  ALLSET_REFERENCE_SRC must point at the reference's `src/` directory."""
import importlib.util
import os
import sys


def load_reference(module_name: str):
    src = os.environ.get('ALLSET_REFERENCE_SRC')
    if not src or not os.path.isfile(os.path.join(src, module_name + '.py')):
        raise ImportError('set ALLSET_REFERENCE_SRC to the AllSet reference `src/` directory (looking for %s.py)' % module_name)
    private = '_allset_reference_' + module_name
    if private in sys.modules:
        return sys.modules[private]
    spec = importlib.util.spec_from_file_location(private, os.path.join(src, module_name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[private] = mod
    spec.loader.exec_module(mod)
    return mod


def public_names(mod):
    return {k: v for k, v in vars(mod).items() if not k.startswith('__')}
