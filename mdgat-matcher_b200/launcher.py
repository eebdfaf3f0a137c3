"""Runs a reference script (test.py, test_registration_metric.py, train.py) UNCHANGED against
the drop-in modules.

    python -m mdgat_matcher_b200.launcher /path/to/MDGAT-matcher/test.py --resume_model ... [script args]

The reference's `models/` directory is a namespace package and the script directory is
sys.path[0], so PYTHONPATH alone cannot override `from models.mdgat import MDGAT`
(SURVEY.md section 8b). The launcher therefore pre-registers the drop-in modules as
sys.modules['models.mdgat'] / ['models.superglue'] (the import system consults sys.modules
first), provides stub modules for the visualisation-only dependencies the scripts import at
top level when they are not installed (open3d, tensorboardX), and then executes the script
with runpy under its own name and argv.
"""
import importlib
import os
import runpy
import sys
import types


def _stub(name, attrs=None):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs or {})
    mod.__stub__ = True
    sys.modules[name] = mod
    return mod


def install_optional_stubs():
    """open3d is used only for visualisation (load_data.py:249-255, utils_test.py:149-403),
    tensorboardX only for scalar logging (train.py:144,309-310)."""
    try:
        importlib.import_module('open3d')
    except Exception:
        class _Unavailable:
            def __getattr__(self, item):
                raise RuntimeError('open3d is not installed; visualisation is unavailable under the launcher')
        o3d = _stub('open3d')
        o3d.geometry = _Unavailable()
        o3d.utility = _Unavailable()
        o3d.visualization = _Unavailable()
    try:
        importlib.import_module('tensorboardX')
    except Exception:
        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def close(self):
                pass
        _stub('tensorboardX', {'SummaryWriter': SummaryWriter})


def register_dropin_modules():
    """After this, `from models.mdgat import MDGAT` and `from models.superglue import SuperGlue`
    resolve to the B200 implementations, whatever is on sys.path."""
    from mdgat_matcher_b200.models import mdgat as _mdgat, superglue as _superglue
    pkg = sys.modules.get('models')
    if pkg is None or not getattr(pkg, '__dropin__', False):
        pkg = types.ModuleType('models')
        pkg.__path__ = []            # a package, with nothing else importable below it by default
        pkg.__dropin__ = True
        sys.modules['models'] = pkg
    sys.modules['models.mdgat'] = _mdgat
    sys.modules['models.superglue'] = _superglue
    pkg.mdgat, pkg.superglue = _mdgat, _superglue
    return pkg


def run_script(path, argv):
    path = os.path.abspath(path)
    install_optional_stubs()
    pkg = register_dropin_modules()
    script_dir = os.path.dirname(path)
    # the scripts also import load_data / utils.utils_test from their own directory
    pkg.__path__ = [os.path.join(script_dir, 'models')]       # models.pointnet.* stay reachable
    sys.path.insert(0, script_dir)
    sys.dont_write_bytecode = True
    old_argv = sys.argv
    sys.argv = [path] + list(argv)
    try:
        return runpy.run_path(path, run_name='__main__')
    finally:
        sys.argv = old_argv


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    run_script(sys.argv[1], sys.argv[2:])


if __name__ == '__main__':
    main()
