import sys, time, logging
sys.path.insert(0, '/root/repo')
from robotic_manipulator_rloa_b200 import ManipulatorFramework
mf = ManipulatorFramework()
t0 = time.time()
mf.run_demo_testing('xarm6_testing')
print('xarm6 demo testing done in', round(time.time() - t0, 1), 's; results:', len(mf.last_test_results),
      'wins', sum(1 for r in mf.last_test_results if r[0]))
