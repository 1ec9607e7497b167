"""
CPU tests of the host-side mirror of the reference's public surface: names, defaults,
validation and error behaviour follow rl_framework.py / environment.py / naf_algorithm.py of the reference, and the
device path refuses to run without librloa_b200.so + a CUDA device instead of falling back to the CPU.
"""
import inspect

import pytest
import torch

import robotic_manipulator_rloa_b200 as pkg
from robotic_manipulator_rloa_b200 import _native
from robotic_manipulator_rloa_b200.environment.environment import EnvironmentConfiguration
from robotic_manipulator_rloa_b200.environment.robot_model import ModelError, load_manipulator
from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent
from robotic_manipulator_rloa_b200.naf_components.naf_neural_network import NAF
from robotic_manipulator_rloa_b200.rl_framework import HyperParameters, ManipulatorFramework
from robotic_manipulator_rloa_b200.utils import exceptions as ex

GOOD = dict(endeffector_index=13, fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
            target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
            initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0], initial_positions_variation_range=[0, 0, .5, .5, .5, .5],
            max_force=200., visualize=False)


def test_public_names_and_signatures():
    assert pkg.ManipulatorFramework is ManipulatorFramework
    mf = ManipulatorFramework
    for name in ('initialize_environment', 'initialize_naf_agent', 'run_training', 'test_trained_model',
                 'load_pretrained_parameters_from_episode', 'load_pretrained_parameters_from_weights_file',
                 'set_hyperparameter', 'get_environment_configuration', 'get_nafagent_configuration',
                 'delete_environment', 'delete_naf_agent', 'run_demo_training', 'run_demo_testing',
                 'plot_training_rewards', 'set_log_level', 'get_required_hyperparameters'):
        assert callable(getattr(mf, name)), name
    sig = inspect.signature(mf.initialize_naf_agent)
    assert sig.parameters['checkpoint_frequency'].default == 500 and sig.parameters['seed'].default == 0
    sig = inspect.signature(mf.run_training)
    assert sig.parameters['frames'].default == 500 and sig.parameters['verbose'].default is True
    sig = inspect.signature(mf.initialize_environment)
    assert sig.parameters['max_force'].default == 200. and sig.parameters['visualize'].default is True
    assert sig.parameters['n_envs'].default == 1                                  # the additive knob keeps 1-env behaviour
    # NAFAgent constructor keyword list (naf_algorithm.py:27-41)
    assert list(inspect.signature(NAFAgent.__init__).parameters)[1:] == [
        'environment', 'state_size', 'action_size', 'layer_size', 'batch_size', 'buffer_size', 'learning_rate', 'tau',
        'gamma', 'update_freq', 'num_updates', 'checkpoint_frequency', 'device', 'seed']
    assert NAFAgent.MODEL_PATH == 'model.p'


def test_hyperparameter_defaults_and_validation():
    hp = HyperParameters()
    assert (hp.buffer_size, hp.batch_size, hp.gamma, hp.tau, hp.learning_rate, hp.update_freq, hp.num_updates) == \
        (100000, 128, 0.99, 0.001, 0.001, 1, 1)                                   # rl_framework.py:68-74
    mf = ManipulatorFramework()
    mf.set_hyperparameter('batch_size', 1024)
    mf.set_hyperparameter('GAMMA', 0.5)
    for name, bad in [('buffer_size', -1), ('batch_size', 1.5), ('gamma', 1.0), ('tau', 2), ('learning_rate', 0),
                      ('update_freq', 0), ('nonsense', 1)]:
        with pytest.raises(ex.InvalidHyperParameter):
            mf.set_hyperparameter(name, bad)


def test_framework_state_errors():
    mf = ManipulatorFramework()
    with pytest.raises(ex.EnvironmentNotInitialized):
        mf.initialize_naf_agent()
    with pytest.raises(ex.ConfigurationIncomplete):
        mf.run_training(1, 1)
    with pytest.raises(ex.ConfigurationIncomplete):
        mf.test_trained_model(1, 1)
    with pytest.raises(ex.EnvironmentNotInitialized):          # checked before the agent, as rl_framework.py:267-271
        mf.load_pretrained_parameters_from_episode(1)
    mf.get_environment_configuration()                         # logs an error, does not raise (rl_framework.py:280-283)
    assert str(ex.MissingWeightsFile()) == 'MissingWeightsFile: The weight file provided does not exist'


# the reference validates types only (environment.py:19-187); same cases as its own parametrisation
# (tests/robotic_manipulator_rloa/environment/test_environment.py:37-53)
@pytest.mark.parametrize('field,bad', [('endeffector_index', 'wrong_type'), ('endeffector_index', 1.5),
                                       ('fixed_joints', 'wrong_type'), ('fixed_joints', ['wrong_type']),
                                       ('involved_joints', 'wrong_type'), ('involved_joints', ['wrong_type']),
                                       ('target_position', 'wrong_type'), ('target_position', ['wrong_type']),
                                       ('obstacle_position', 'wrong_type'), ('obstacle_position', ['wrong_type']),
                                       ('initial_joint_positions', 'wrong_type'), ('initial_joint_positions', ['wrong_type']),
                                       ('initial_positions_variation_range', 'wrong_type'),
                                       ('initial_positions_variation_range', ['wrong_type']),
                                       ('max_force', 'wrong_type'), ('visualize', 'wrong_type')])
def test_environment_configuration_rejects_invalid_parameters(field, bad):
    EnvironmentConfiguration(**GOOD)
    with pytest.raises(ex.InvalidEnvironmentParameter):
        EnvironmentConfiguration(**{**GOOD, field: bad})


def test_model_files():
    kuka = load_manipulator('kuka_iiwa/kuka_with_gripper2.sdf')
    assert kuka.nl == 14 and len(kuka.joint_names) == 14
    panda = load_manipulator('franka_panda/panda.urdf')
    assert panda.nl == 12 and sorted(set(int(t) for t in panda.jtype)) == [0, 1, 2]   # fixed, revolute, prismatic
    xarm = load_manipulator('xarm/xarm6_with_gripper.urdf')                       # the reference's second demo asset
    assert xarm.nl == 14 and int(xarm.jtype[0]) == 0 and int(xarm.jtype[7]) == 0   # world joint and gripper mount fixed
    with pytest.raises(ModelError):
        load_manipulator('no/such/file.urdf')
    mf = ManipulatorFramework()
    with pytest.raises(ex.InvalidManipulatorFile):
        mf.initialize_environment('robot.txt', **{k: v for k, v in GOOD.items() if k not in ('visualize',)}, visualize=False)


def test_no_cpu_fallback():
    """The device path fails loudly without CUDA: parameters on the CPU are refused, nothing routes through torch."""
    net = NAF(21, 6, 256, 0, torch.device('cpu'))
    assert list(net.state_dict().keys())[:4] == ['input_layer.weight', 'input_layer.bias', 'bn1.weight', 'bn1.bias']
    assert len(net.state_dict()) == 20                                             # the .p checkpoint layout
    with pytest.raises(_native.NativeLibraryError):
        net(torch.zeros(2, 21))
    agent = NAFAgent(None, 21, 6, 256, 128, 1000, 1e-3, 1e-3, 0.99, 1, 1, 500, torch.device('cpu'), 0)
    with pytest.raises(_native.NativeLibraryError):
        agent.act_batch(torch.zeros(2, 21))
    with pytest.raises(_native.NativeLibraryError):
        agent.learn(tuple(torch.zeros(4, k) for k in (21, 6, 1, 21, 1)))
