"""
ManipulatorFramework — the user-facing facade, drop-in for
/root/reference/robotic_manipulator_rloa/rl_framework.py: same methods, argument meaning, defaults,
validation rules and exceptions.  Everything below it (Environment, NAFAgent, ReplayBuffer, NAF) runs on the
B200 through librloa_b200.so.  Additive knobs only: ``n_envs`` / ``device`` on ``initialize_environment``
(defaults reproduce the single-env behaviour), ``set_trunk_mode``.

Host-side artefacts (logger, hyper-parameter store, reward plot, demos) are thin Python: they are not on the
hot path.
"""
from __future__ import annotations

import json
import logging
import os
from dataclasses import dataclass
from typing import List, Optional, Union

import numpy as np
import torch

from .environment.environment import Environment, EnvironmentConfiguration
from .environment.robot_model import DATA_PATH
from .naf_components.naf_algorithm import NAFAgent
from .utils.exceptions import (ConfigurationIncomplete, EnvironmentNotInitialized, InvalidHyperParameter,
                               InvalidNAFAgentParameter, NAFAgentNotInitialized)
from .utils.logger import Logger, get_global_logger

logger = get_global_logger()
Logger.set_logger_setup()          # like the reference, importing the package installs the logger (B.13)

_PKG_DIR = os.path.dirname(os.path.realpath(__file__))


@dataclass
class HyperParameters:
    buffer_size: int = 100000
    batch_size: int = 128
    gamma: float = 0.99
    tau: float = 0.001
    learning_rate: float = 0.001
    update_freq: int = 1
    num_updates: int = 1


def _is_pos_int(v) -> bool:
    return isinstance(v, int) and v > 0


# accepted spellings -> (attribute, validator, message); same names / ranges / messages as rl_framework.py:179-230
_HYPERPARAMETER_RULES = [
    (('buffer_size', 'buffersize', 'BUFFER_SIZE', 'BUFFERSIZE'), 'buffer_size', _is_pos_int,
     'Buffer Size is not an int or has a value lower than 0'),
    (('batch_size', 'batchsize', 'BATCH_SIZE', 'BATCHSIZE'), 'batch_size', _is_pos_int,
     'Batch Size is not an int or has a value lower than 0'),
    (('gamma', 'GAMMA'), 'gamma', lambda v: isinstance(v, (int, float)) and 0 < v < 1,
     'Gamma is not a float or its value is out of range (0, 1)'),
    (('tau', 'TAU'), 'tau', lambda v: isinstance(v, (int, float)) and 0 <= v <= 1,
     'Tau is not a float or its value is out of range [0, 1]'),
    (('learning_rate', 'learningrate', 'LEARNING_RATE', 'LEARNINGRATE'), 'learning_rate',
     lambda v: isinstance(v, (int, float)) and v > 0, 'Learning Rate is not a float or has a value lower than 0'),
    (('update_freq', 'updatefreq', 'UPDATE_FREQ', 'UPDATEFREQ'), 'update_freq', _is_pos_int,
     'Update Frequency is not an int or has a value lower than 0'),
    (('num_update', 'numupdate', 'NUMUPDATE', 'NUM_UPDATE'), 'num_updates', _is_pos_int,
     'Buffer Size is not an int or has a value lower than 0'),
]

_DEMOS = {
    'kuka': dict(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                 fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                 target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                 initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0]),
    'xarm6': dict(manipulator_file='xarm/xarm6_with_gripper.urdf', endeffector_index=12,
                  fixed_joints=[0, 7, 8, 9, 10, 11, 12, 13], involved_joints=[1, 2, 3, 4, 5, 6],
                  target_position=[0.3, 0.47, 0.61], obstacle_position=[0.25, 0.27, 0.5],
                  initial_joint_positions=[0., 1., 0., -2.3, 0., 0., 0.],
                  initial_positions_variation_range=[0, 0, 0, 0.3, 1, 1, 1]),
}


class ManipulatorFramework:

    def __init__(self) -> None:
        self.env: Union[Environment, None] = None
        self.naf_agent: Union[NAFAgent, None] = None
        self._hyperparameters: Union[HyperParameters, None] = None
        self._initialize_hyperparameters()
        logger.info('The Framework has been initialized with the default hyperparameters configuration')
        logger.debug('* Custom hyperparameters can be set via the set_hyperparameter() method')
        logger.debug('* All the required hyperparameters can be printed via the get_required_hyperparameters() method')
        logger.debug('* Load a manipulator via the initialize_environment() method to start with the training '
                     'configuration')

    def _initialize_hyperparameters(self) -> None:
        self._hyperparameters = HyperParameters(buffer_size=100000, batch_size=128, gamma=0.99, tau=0.001,
                                                learning_rate=0.001, update_freq=1, num_updates=1)

    # ---- logging helpers ------------------------------------------------------------------------
    @staticmethod
    def set_log_level(log_level: int) -> None:
        names = {10: 'DEBUG', 20: 'INFO', 30: 'WARNING', 40: 'ERROR', 50: 'CRITICAL'}
        if isinstance(log_level, int) and log_level in names:
            logger.info(f'Log Level has been set to {logger.level} ({names[logger.level]})')
            logger.setLevel(log_level)
        else:
            logger.error(f'The Log level provided is invalid, so the previous Log Level is maintained ({logger.level}))')
            logger.error('Valid values: 10 (DEBUG), 20 (INFO), 30 (WARNING), 40 (ERROR), 50 (CRITICAL)')

    @staticmethod
    def get_required_hyperparameters() -> None:
        if logger.level > 10:
            logger.error('get_required_hyperparameters() only shows information for DEBUG log level. '
                         'Try running this method after setting the log level to DEBUG by calling '
                         'set_log_level(10) class method')
            return
        medium = ('https://medium.com/towards-data-science/applied-reinforcement-learning-v-normalized'
                  '-advantage-function-naf-for-continuous-control-62ad143d3095')
        info = {
            'Buffer Size': 'https://www.tensorflow.org/agents/tutorials/5_replay_buffers_tutorial?hl=es-419',
            'Batch Size': 'https://www.kaggle.com/general/276990',
            'Gamma (discount factor)': 'https://arxiv.org/pdf/2007.02040.pdf',
            'Tau': 'https://arxiv.org/abs/1603.00748',
            'Learning Rate': 'https://machinelearningmastery.com/understand-the-dynamics-of-learning-rate-on-deep'
                             '-learning-neural-networks/',
            'Update Frequency': medium,
            'Number of Updates': medium}
        logger.debug('Required Hyperparameters:')
        for name, link in info.items():
            logger.debug('{:<25} (see {:<10})'.format(name, link))

    @staticmethod
    def plot_training_rewards(episode: int, mean_range: int = 50) -> None:
        """Mean episode reward per window of ``mean_range`` episodes from checkpoints/{episode}/scores.txt."""
        try:
            with open(f'checkpoints/{episode}/scores.txt', 'r') as f:
                scores = json.loads(f.read())
        except FileNotFoundError as err:
            logger.error(f'File "scores.txt" located in checkpoints/{episode}/ folder was not found')
            raise err
        rewards = [result[0] for result in scores.values()]
        windows = [rewards[i:i + mean_range] for i in range(0, len(rewards) - mean_range + 1, mean_range)]
        values_to_plot = [sum(w) / len(w) for w in windows]
        import matplotlib.pyplot as plt          # optional dependency, only needed here
        plt.set_loglevel('critical')
        logging.getLogger('PIL').setLevel(logging.WARNING)
        plt.plot(range(len(values_to_plot)), values_to_plot)
        plt.show()

    # ---- hyper-parameters -----------------------------------------------------------------------
    def set_hyperparameter(self, hyperparameter: str, value: Union[float, int]) -> None:
        for names, attr, ok, message in _HYPERPARAMETER_RULES:
            if hyperparameter in names:
                if not ok(value):
                    raise InvalidHyperParameter(message)
                setattr(self._hyperparameters, attr, value)
                logger.info(f'Hyperparameter {hyperparameter} has been set to {value}')
                return
        raise InvalidHyperParameter(
            'The hyperparameter name passed as parameter is not valid. Valid hyperparameters are: '
            '["buffer_size", "batch_size", "gamma", "tau", "learning_rate", "update_freq", "num_update"]')

    # ---- weights --------------------------------------------------------------------------------
    def load_pretrained_parameters_from_weights_file(self, parameters_file_path: str) -> None:
        if not self.env:
            raise EnvironmentNotInitialized
        if not self.naf_agent:
            raise NAFAgentNotInitialized
        self.naf_agent.initialize_pretrained_agent_from_weights_file(parameters_file_path)

    def load_pretrained_parameters_from_episode(self, episode: int) -> None:
        if not self.env:
            raise EnvironmentNotInitialized
        if not self.naf_agent:
            raise NAFAgentNotInitialized
        self.naf_agent.initialize_pretrained_agent_from_episode(episode)

    # ---- configuration dumps --------------------------------------------------------------------
    def get_environment_configuration(self) -> None:
        if not self.env:
            logger.error("Environment is not initialized yet, can't show configuration")
            return
        e = self.env
        logger.info('Environment Configuration:')
        for label, val in [('Manipulator File:                   ', e.manipulator_file),
                           ('End Effector index:                 ', e.endeffector_index),
                           ('List of fixed Joints:               ', e.fixed_joints),
                           ('List of Joints involved in training:', e.involved_joints),
                           ('Position of the Target:             ', e.target_pos),
                           ('Position of the Obstacle:           ', e.obstacle_pos),
                           ('Initial position of joints:         ', e.initial_joint_positions),
                           ('Initial variation range of joints:  ', e.initial_positions_variation_range),
                           ('Max Force to be applied on joints:  ', e.max_force),
                           ('Visualize mode:                     ', e.visualize),
                           ('Instance of the Environment:        ', e)]:
            logger.info(f'* {label} {val}')

    def get_nafagent_configuration(self) -> None:
        if not self.naf_agent:
            logger.error("NAFAgent is not initialized yet, can't show configuration")
            return
        a = self.naf_agent
        logger.info('NAFAgent Configuration:')
        for label, val in [('Environment Instance:                   ', a.environment),
                           ('State Size:                             ', a.state_size),
                           ('Action Size:                            ', a.action_size),
                           ('Size of layers of the Neural Network:   ', a.layer_size),
                           ('Batch Size:                             ', a.batch_size),
                           ('Buffer Size:                            ', a.buffer_size),
                           ('Learning Rate:                          ', a.learning_rate),
                           ('Tau:                                    ', a.tau),
                           ('Gamma:                                  ', a.gamma),
                           ('Update Frequency:                       ', a.update_freq),
                           ('Number of Updates:                      ', a.num_updates),
                           ('Checkpoint frequency:                   ', a.checkpoint_frequency),
                           ('Device:                                 ', a.device)]:
            logger.info(f'* {label} {val}')

    # ---- rollouts -------------------------------------------------------------------------------
    def test_trained_model(self, n_episodes: int, frames: int) -> None:
        """``n_episodes`` test episodes of at most ``frames`` steps (rl_framework.py:319-367), spread over the
        environment's ``n_envs`` arms; results are logged in the reference's format and kept in
        ``self.last_test_results`` as (completed, frame) tuples."""
        if not self.naf_agent or not self.env:
            raise ConfigurationIncomplete
        env, agent = self.env, self.naf_agent
        n, dev = env.n_envs, agent.device
        results, num_collisions = list(), 0
        if n == 1:
            for ep in range(n_episodes):
                state = env.reset()
                for frame in range(frames):
                    action = agent.act(state)
                    next_state, reward, done = env.step(action)
                    state = next_state
                    if done:
                        results.append((reward == 250, frame))
                        num_collisions += 0 if reward == 250 else 1
                        break
                    if frame == frames - 1 and not done:
                        results.append((False, frame))
                        break
                logger.info('Test Episode number {ep} completed\n'.format(ep=ep + 1))
        else:
            # waves of n episodes: every env runs one episode; finished envs idle (active = 0) until the wave ends
            remaining = n_episodes
            f32 = dict(dtype=torch.float32, device=dev)
            obs = torch.empty(n, env.sim.obs_size, **f32)
            reward = torch.zeros(n, **f32)
            done = torch.zeros(n, dtype=torch.uint8, device=dev)
            while remaining > 0:
                wave = min(n, remaining)
                active = torch.zeros(n, dtype=torch.uint8, device=dev)
                active[:wave] = 1
                env.reset_batch(mask=active, obs=obs)
                end_frame = torch.full((n,), -1, dtype=torch.int32, device=dev)
                end_reward = torch.zeros(n, **f32)
                for frame in range(frames):
                    actions = agent.act_batch(obs)
                    env.sim.step(actions, active=active, out=(obs, reward, done))
                    fin = (done != 0) & (active != 0)
                    end_frame = torch.where(fin, torch.full_like(end_frame, frame), end_frame)
                    end_reward = torch.where(fin, reward, end_reward)
                    active = active & (~fin).to(torch.uint8)
                    if frame % 16 == 15 and int(active.sum().item()) == 0:
                        break
                ef, er = end_frame[:wave].cpu().numpy(), end_reward[:wave].cpu().numpy()
                for k in range(wave):
                    if ef[k] < 0:
                        results.append((False, frames - 1))
                    else:
                        results.append((bool(er[k] == 250), int(ef[k])))
                        num_collisions += 0 if er[k] == 250 else 1
                    logger.info('Test Episode number {ep} completed\n'.format(ep=len(results)))
                remaining -= wave
        self.last_test_results = results
        logger.info('RESULTS OF THE TEST:')
        for i, result in enumerate(results):
            logger.info(f'Results of Iteration {i + 1}: COMPLETED: {result[0]}. FRAMES: {result[1]}')
        wins = [res[0] for res in results].count(True)
        logger.info(f'Number of successful executions: {wins}/{len(results)}  ({(wins / len(results)) * 100}%)')
        logger.info(f'Average number of frames required to complete an episode: '
                    f'{np.mean(np.array([res[1] for res in results if res[0]]))}')
        logger.info(f'Number of episodes terminated because of collisions: {num_collisions}')

    # ---- construction ---------------------------------------------------------------------------
    def initialize_environment(self, manipulator_file: str, endeffector_index: int, fixed_joints: List[int],
                               involved_joints: List[int], target_position: List[float],
                               obstacle_position: List[float], initial_joint_positions: List[float] = None,
                               initial_positions_variation_range: List[float] = None, max_force: float = 200.,
                               visualize: bool = True, n_envs: int = 1, device: Optional[torch.device] = None) -> None:
        logger.debug('Initializing Pybullet Environment...')
        environment_config = EnvironmentConfiguration(
            endeffector_index=endeffector_index, fixed_joints=fixed_joints, involved_joints=involved_joints,
            target_position=target_position, obstacle_position=obstacle_position,
            initial_joint_positions=initial_joint_positions,
            initial_positions_variation_range=initial_positions_variation_range, max_force=max_force,
            visualize=visualize)
        kwargs = {}
        if n_envs != 1:
            kwargs['n_envs'] = n_envs
        if device is not None:
            kwargs['device'] = device
        self.env = Environment(manipulator_file=manipulator_file, environment_config=environment_config, **kwargs)
        logger.info('Pybullet Environment successfully initialized')
        logger.debug('* The NAF Agent can now be initialized via the initialize_naf_agent() method')

    def delete_environment(self) -> None:
        if not self.env:
            logger.error('No existing instance of Environment found')
            return
        if hasattr(self.env, 'close'):
            self.env.close()
        self.env = None
        logger.info('Environment instance has been successfully removed')

    def initialize_naf_agent(self, checkpoint_frequency: int = 500, seed: int = 0) -> None:
        if not self.env:
            raise EnvironmentNotInitialized
        if not isinstance(checkpoint_frequency, int) or not isinstance(seed, int):
            raise InvalidNAFAgentParameter('Checkpoint Frequency or Seed received is not an integer')
        logger.debug('Initializing NAF Agent...')
        device = getattr(self.env, 'device', None)
        if not isinstance(device, torch.device):
            device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
        hp = self._hyperparameters
        self.naf_agent = NAFAgent(environment=self.env, state_size=self.env.observation_space.shape[0],
                                  action_size=self.env.action_space.shape[0], layer_size=256,
                                  batch_size=hp.batch_size, buffer_size=hp.buffer_size,
                                  learning_rate=hp.learning_rate, tau=hp.tau, gamma=hp.gamma,
                                  update_freq=hp.update_freq, num_updates=hp.num_updates,
                                  checkpoint_frequency=checkpoint_frequency, device=device, seed=seed)
        logger.info('NAF Agent successfully initialized')
        logger.debug('* The Robotic Manipulator training can now be launched via the run_training() method')

    def delete_naf_agent(self) -> None:
        if not self.naf_agent:
            logger.error('No existing instance of NAFAgent found')
            return
        self.naf_agent = None
        logger.info('NAFAgent instance has been successfully removed')

    def run_training(self, episodes: int, frames: Optional[int] = 500, verbose: bool = True):
        if not self.naf_agent or not self.env:
            raise ConfigurationIncomplete
        return self.naf_agent.run(frames, episodes, verbose)

    # ---- demos (rl_framework.py:503-699) ---------------------------------------------------------
    def _clear_for_demo(self) -> bool:
        if self.env:
            if input('Environment instance found. Overwrite? [Y/n] ').lower() != 'y':
                logger.info('Demo could not run due to the presence of a user-configured Environment instance')
                return False
            self.delete_environment()
        if self.naf_agent:
            if input('NAFAgent instance found. Overwrite? [Y/n] ').lower() != 'y':
                logger.info('Demo could not run due to the presence of a user-configured NAFAgent instance')
                return False
            self.delete_naf_agent()
        return True

    def run_demo_training(self, demo_type: str, verbose: bool = False) -> None:
        logger.warning('Both the demo testing and the demo training are executed with the Log level '
                       'set to DEBUG, so that the framework can be understood at a low level.')
        old_level = logger.level
        logger.setLevel(10)
        if not self._clear_for_demo():
            return
        if demo_type in ('kuka_training', 'xarm6_training'):
            cfg = dict(_DEMOS[demo_type.split('_')[0]])
            if demo_type == 'kuka_training':
                cfg['initial_positions_variation_range'] = [0, 0, 0, 0, 0, 0]
            else:
                cfg['manipulator_file'] = os.path.join(DATA_PATH, cfg['manipulator_file'])
            logger.info('Initializing demo Environment instance...')
            self.initialize_environment(visualize=True, **cfg)
            logger.info('Initializing demo NAFAgent instance...')
            self.initialize_naf_agent()
            logger.info('Running training for 20 episodes. Do not expect good results, '
                        'this is just a demo of the training configuration process')
            self.run_training(20, 400, verbose=verbose)
            self.delete_environment()
            self.delete_naf_agent()
        else:
            logger.error('Incorrect demo type!')
        logger.setLevel(old_level)
        logger.warning('Log level has been reset to its original value')

    def run_demo_testing(self, demo_type: str) -> None:
        logger.warning('Both the demo testing and the demo training are executed with the Log level '
                       'set to DEBUG, so that the framework can be understood at a low level.')
        old_level = logger.level
        logger.setLevel(10)
        if not self._clear_for_demo():
            return
        if demo_type in ('kuka_testing', 'xarm6_testing'):
            name = demo_type.split('_')[0]
            cfg = dict(_DEMOS[name])
            cfg['manipulator_file'] = os.path.join(DATA_PATH, cfg['manipulator_file'])
            if name == 'kuka':
                cfg['initial_positions_variation_range'] = [0, 0, .5, .5, .5, .5]
            logger.info('Initializing demo Environment instance...')
            self.initialize_environment(**cfg)
            logger.info('Initializing demo NAFAgent instance...')
            self.initialize_naf_agent()
            logger.info('Loading demo pretrained parameters')
            self.load_pretrained_parameters_from_weights_file(
                _PKG_DIR + f'/naf_components/demo_weights/weights_{name}.p')
            logger.info('Running 50 test episodes...')
            self.test_trained_model(50, 750)
            self.delete_environment()
            self.delete_naf_agent()
        else:
            logger.error('Incorrect demo type!')
        logger.setLevel(old_level)
        logger.warning('Log level has been reset to its original value')
