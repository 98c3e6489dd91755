"""Import alias: ``import truely_b200`` -> the package directory whose name the task fixes
(``truely-real-time-ai-generated-video-detection-framework-for-social-platforms_b200``, not a
valid identifier, hence this shim).  Sub-modules are aliased too so no module is loaded twice."""
import importlib
import sys

PACKAGE_DIR_NAME = "truely-real-time-ai-generated-video-detection-framework-for-social-platforms_b200"
_pkg = importlib.import_module(PACKAGE_DIR_NAME)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(PACKAGE_DIR_NAME + "."):
        sys.modules["truely_b200" + _name[len(PACKAGE_DIR_NAME):]] = _mod
sys.modules["truely_b200"] = _pkg
