"""Import alias: ``import sac_agent_b200`` loads the package that lives in the
``sac-agent_b200/`` directory (a hyphen is not importable)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "sac-agent_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f, _os, _real
