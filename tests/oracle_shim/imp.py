"""Stand-in for the stdlib `imp` module (removed in Python 3.12); the reference's
scenario loader calls `imp.load_source(name, path)`."""
import importlib.util


def load_source(name, path):
    spec = importlib.util.spec_from_file_location(name or "_scn", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
