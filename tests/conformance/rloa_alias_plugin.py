"""pytest plugin (loaded with -p) that lets the REFERENCE's own test-suite import this package under the reference's
name: `robotic_manipulator_rloa[.x.y]` resolves to `robotic_manipulator_rloa_b200[.x.y]`, the `mock` distribution to
`unittest.mock`, and the three third-party modules the reference imports at package level (pybullet, pybullet_data,
matplotlib) to stubs.  Used by tests/test_reference_suite_conformance.py only."""
import importlib
import os
import sys
import types
import unittest.mock

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

sys.modules.setdefault('mock', unittest.mock)
for name in ('pybullet', 'pybullet_data', 'matplotlib', 'matplotlib.pyplot'):
    if name not in sys.modules:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = unittest.mock.MagicMock(name=name)

NEW, OLD = 'robotic_manipulator_rloa_b200', 'robotic_manipulator_rloa'
pkg = importlib.import_module(NEW)
for sub in ('rl_framework', 'environment', 'environment.environment', 'naf_components', 'naf_components.naf_algorithm',
            'naf_components.naf_neural_network', 'utils', 'utils.collision_detector', 'utils.exceptions', 'utils.logger',
            'utils.replay_buffer'):
    importlib.import_module(f'{NEW}.{sub}')
for name, mod in list(sys.modules.items()):
    if name == NEW or name.startswith(NEW + '.'):
        sys.modules[OLD + name[len(NEW):]] = mod
