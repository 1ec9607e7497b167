"""
Environment — drop-in for /root/reference/robotic_manipulator_rloa/environment/environment.py, with the
PyBullet world replaced by the batched sm_100a simulator (environment/simulator.py, rloa_sim_*).

Single-env use keeps the reference contract: ``reset() -> ndarray[S] float64``,
``step(a) -> (ndarray[S] float64, reward, done int)``.  ``n_envs > 1`` (additive knob) exposes the same
step / reset over device tensors for the vectorised training loop.
No PyBullet call and no CPU physics fallback exist on this path.
"""
from __future__ import annotations

import random
from typing import List, Optional, Tuple

import numpy as np
import torch
from numpy.typing import NDArray

from ..utils.collision_detector import CollisionDetector, CollisionObject
from ..utils.exceptions import InvalidEnvironmentParameter, InvalidManipulatorFile
from ..utils.logger import get_global_logger
from .robot_model import ModelError, load_manipulator
from .simulator import BatchedSimulator

logger = get_global_logger()


def _require_list_of(value, types, what: str, item_what: str):
    if not isinstance(value, list):
        raise InvalidEnvironmentParameter(f'{what} received is not a list')
    for item in value:
        if not isinstance(item, types):
            raise InvalidEnvironmentParameter(f'An item inside the {what} list is not {item_what}')
    return value


class EnvironmentConfiguration:
    """Type validation of the Environment parameters (reference environment.py:19-187: same checks, same
    InvalidEnvironmentParameter messages), expressed through one helper."""

    def __init__(self, endeffector_index: int, fixed_joints: List[int], involved_joints: List[int],
                 target_position: List[float], obstacle_position: List[float],
                 initial_joint_positions: List[float] = None, initial_positions_variation_range: List[float] = None,
                 max_force: float = 200., visualize: bool = True):
        if not isinstance(endeffector_index, int):
            raise InvalidEnvironmentParameter('End Effector index received is not an integer')
        self.endeffector_index = endeffector_index
        self.fixed_joints = _require_list_of(fixed_joints, int, 'Fixed Joints', 'an integer')
        self.involved_joints = _require_list_of(involved_joints, int, 'Involved Joints', 'an integer')
        self.target_position = _require_list_of(target_position, (int, float), 'Target Position', 'a float')
        self.obstacle_position = _require_list_of(obstacle_position, (int, float), 'Obstacle Position', 'a float')
        self.initial_joint_positions = None if initial_joint_positions is None else _require_list_of(
            initial_joint_positions, (int, float), 'Initial Joint Positions', 'a float')
        self.initial_positions_variation_range = None if initial_positions_variation_range is None else \
            _require_list_of(initial_positions_variation_range, (float, int), 'Initial Positions Variation Range',
                             'a float')
        if not isinstance(max_force, (int, float)):
            raise InvalidEnvironmentParameter('Maximum Force value received is not a float')
        self.max_force = max_force
        if not isinstance(visualize, bool):
            raise InvalidEnvironmentParameter('Visualize value received is not a boolean')
        self.visualize = visualize


class Environment:

    def __init__(self, manipulator_file: str, environment_config: EnvironmentConfiguration, n_envs: int = 1,
                 device: Optional[torch.device] = None, seed: int = 0):
        self.manipulator_file = manipulator_file
        self.visualize = environment_config.visualize
        if self.visualize:
            logger.warning('visualize=True: there is no GUI on the B200 device path, running headless')
        self.physics_client = 0                      # the reference stores PyBullet's client id here
        self.target_pos = environment_config.target_position
        self.obstacle_pos = environment_config.obstacle_position
        self.max_force = environment_config.max_force
        self.initial_joint_positions = environment_config.initial_joint_positions
        self.initial_positions_variation_range = environment_config.initial_positions_variation_range
        self.endeffector_index = environment_config.endeffector_index
        self.fixed_joints = environment_config.fixed_joints
        self.involved_joints = environment_config.involved_joints
        self.n_envs = int(n_envs)

        logger.debug(f'Loading URDF/SDF file {manipulator_file} for Robot Manipulator...')
        if not isinstance(manipulator_file, str):
            raise InvalidManipulatorFile('The filename provided is not a string')
        if not manipulator_file.endswith(('.urdf', '.sdf')):
            raise InvalidManipulatorFile('The file extension is neither .sdf nor .urdf')
        try:
            self.model = load_manipulator(manipulator_file)
        except ModelError as err:
            logger.critical(err)
            raise InvalidManipulatorFile
        self.manipulator_uid = 0
        self.num_joints = self.model.nl
        logger.debug(f'Robot Manipulator URDF/SDF file {manipulator_file} has been successfully loaded. '
                     f'The Robot Manipulator has {self.num_joints} joints, and its joints, '
                     f'together with the information of each, are:')
        self.print_table([(i, self.model.joint_names[i], float(self.model.upper[i]), float(self.model.lower[i]),
                           tuple(float(x) for x in self.model.joint_axis_link[i])) for i in range(self.num_joints)])

        self.device = torch.device(device if device is not None else 'cuda:0')
        self.sim = BatchedSimulator(self.model, self.n_envs, self.endeffector_index, self.involved_joints,
                                    self.fixed_joints, max_force=float(self.max_force), device=self.device)
        self.sim.endeffector_index = self.endeffector_index
        self.sim.set_task(self.target_pos, self.obstacle_pos)
        self.obstacle, self.target = 'obstacle', 'target'
        logger.debug(f'Both the obstacle and the target object have been generated in positions {self.obstacle_pos} '
                     f'and {self.target_pos} respectively')
        self._observation_space = np.zeros((9 + 2 * len(self.involved_joints),))
        self._action_space = np.zeros((len(self.involved_joints),))
        # one process per GPU: every rank draws its own start poses
        from ..utils import distributed as rdist
        rank, world = rdist.world()
        self.seed = int(seed) if world == 1 else rdist.rank_seed(int(seed), rank, 2)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed((self.seed + 0x5EED) & 0x7FFFFFFFFFFFFFFF)
        n_init = self._n_init()
        pos = list(self.initial_joint_positions[:n_init]) if self.initial_joint_positions else [0.0] * n_init
        var = (list(self.initial_positions_variation_range[:n_init]) if self.initial_positions_variation_range
               else [0.0] * n_init)
        self._d_pos = torch.tensor(pos, dtype=torch.float32, device=self.device)
        self._d_var = torch.tensor(var, dtype=torch.float32, device=self.device)
        # pinned staging for the single-env host API
        self._h_action = torch.zeros(self.n_envs, len(self.involved_joints), dtype=torch.float32).pin_memory()
        self._d_action = torch.zeros(self.n_envs, len(self.involved_joints), dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------------------------------
    # start poses (reference environment.py:284-293)
    def _n_init(self) -> int:
        if not self.initial_joint_positions and not self.initial_positions_variation_range:
            return self.num_joints
        if self.initial_joint_positions:
            if self.initial_positions_variation_range:
                return min(len(self.initial_joint_positions), len(self.initial_positions_variation_range))
            return len(self.initial_joint_positions)
        return len(self.initial_positions_variation_range)

    def _initial_state_host(self) -> List[float]:
        """One start pose drawn with the global Python RNG, exactly like the reference does."""
        if not self.initial_joint_positions and not self.initial_positions_variation_range:
            return [0 for _ in range(self.num_joints)]
        if self.initial_joint_positions:
            if self.initial_positions_variation_range:
                return [random.uniform(pos - var, pos + var) for pos, var
                        in zip(self.initial_joint_positions, self.initial_positions_variation_range)]
            return list(self.initial_joint_positions)
        return [random.uniform(0 - var, 0 + var) for var in self.initial_positions_variation_range]

    def initial_targets(self) -> torch.Tensor:
        """Start poses for every env, fp32 [n_envs, n_init] on the device (uniform noise from the device RNG)."""
        n_init = self._n_init()
        pos = torch.zeros(n_init, dtype=torch.float32)
        var = torch.zeros(n_init, dtype=torch.float32)
        if self.initial_joint_positions:
            pos = torch.tensor(self.initial_joint_positions[:n_init], dtype=torch.float32)
        if self.initial_positions_variation_range:
            var = torch.tensor(self.initial_positions_variation_range[:n_init], dtype=torch.float32)
        pos, var = pos.to(self.device), var.to(self.device)
        u = torch.rand(self.n_envs, n_init, generator=self._gen, device=self.device, dtype=torch.float32)
        return (pos + var * (2.0 * u - 1.0)).contiguous()

    def begin_reset_masked(self, mask: torch.Tensor, tick: Optional[torch.Tensor] = None) -> None:
        """Lock-step asynchronous reset of the masked envs with start poses drawn on the device
        (environment.py:284-301); one launch, no host work, CUDA-graph capturable."""
        self.sim.begin_reset_random(self._d_pos, self._d_var, self._n_init(), mask=mask, seed=self.seed + 0x5EED,
                                    tick=tick)

    # ------------------------------------------------------------------------------------------
    # batched API (device tensors)
    def reset_batch(self, mask: Optional[torch.Tensor] = None, init_targets: Optional[torch.Tensor] = None,
                    obs: Optional[torch.Tensor] = None) -> torch.Tensor:
        if init_targets is None:
            init_targets = self.initial_targets()
        return self.sim.reset(init_targets, mask=mask, obs=obs)

    def step_batch(self, actions: torch.Tensor, active: Optional[torch.Tensor] = None, out=None):
        return self.sim.step(actions, active=active, out=out)

    def set_task_positions(self, target, obstacle) -> None:
        """Per-env target / obstacle positions ([n_envs,3] or [3])."""
        self.sim.set_task(target, obstacle)

    # ------------------------------------------------------------------------------------------
    # reference API
    def reset(self, verbose: bool = True):
        if verbose: logger.info('Resetting Environment...')
        if self.n_envs == 1:
            init = torch.tensor([self._initial_state_host()], dtype=torch.float32, device=self.device)
            obs = self.sim.reset(init)
            new_state = obs[0].cpu().numpy().astype(float)
        else:
            new_state = self.reset_batch().clone()
        if verbose: logger.info('Environment Reset')
        return new_state

    def step(self, action) -> Tuple[NDArray, float, int]:
        if self.n_envs != 1 or isinstance(action, torch.Tensor) and action.is_cuda:
            obs, reward, done = self.step_batch(action if isinstance(action, torch.Tensor)
                                                else torch.as_tensor(np.asarray(action), dtype=torch.float32,
                                                                     device=self.device))
            return obs, reward, done
        self._h_action[0].copy_(torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(-1)))
        self._d_action.copy_(self._h_action, non_blocking=True)
        obs, reward, done = self.sim.step(self._d_action)
        packed = torch.cat([obs[0], reward[:1], done[:1].float()]).cpu().numpy()
        S = obs.shape[1]
        r = float(packed[S])
        reward_out = int(r) if r in (250.0, -1000.0) else r
        return packed[:S].astype(float), reward_out, int(packed[S + 1])

    def get_state(self) -> NDArray:
        obs = self.sim.observe()
        return obs[0].cpu().numpy().astype(float) if self.n_envs == 1 else obs

    def get_manipulator_obstacle_collisions(self, threshold: float) -> bool:
        _, link_obstacle, _ = self.sim.observe(want_distances=True)
        return bool((link_obstacle[0] < threshold).any().item())

    def get_manipulator_collisions_with_itself(self) -> dict:
        """``{'joint_i': distances from link i to every link except i-1, i, i+1}`` like the reference
        (environment.py:394-412); arrays are [k] for one env and [n_envs][k] tensors for a batch."""
        d = self.sim.self_distances()
        nl = self.num_joints
        out = {}
        for i in range(nl):
            others = [j for j in range(nl) if abs(j - i) > 1]
            sel = d[:, i, others]
            out[f'joint_{i}'] = sel[0].cpu().numpy().astype(float) if self.n_envs == 1 else sel
        return out

    def _self_collision(self) -> bool:
        return any(bool((v < 0).any()) for v in self.get_manipulator_collisions_with_itself().values())

    def get_endeffector_target_collision(self, threshold: float) -> Tuple[bool, NDArray]:
        ee = CollisionObject(body=self.sim, link=self.endeffector_index)
        dist = CollisionDetector(collision_object=ee, obstacle_ids=[self.target]).compute_distances()
        return bool((dist < threshold).any()), dist - threshold

    def is_terminal_state(self, target_threshold: float = 0.05, obstacle_threshold: float = 0.,
                          consider_autocollision: bool = False) -> int:
        if self.get_manipulator_obstacle_collisions(threshold=obstacle_threshold):
            logger.info('Collision detected, terminating episode...')
            return 1
        if self.get_endeffector_target_collision(threshold=target_threshold)[0]:
            logger.info('The goal state has been reached, terminating episode...')
            return 1
        if consider_autocollision and self._self_collision():
            logger.info('Auto-Collision detected, terminating episode...')
            return 1
        return 0

    def get_reward(self, consider_autocollision: bool = False) -> float:
        self_collision = consider_autocollision and self._self_collision()
        hit_target, dist = self.get_endeffector_target_collision(threshold=0.05)
        if hit_target:
            return 250
        if self.get_manipulator_obstacle_collisions(threshold=0) or self_collision:
            return -1000
        return -1 * float(dist[0])

    @staticmethod
    def print_table(data) -> None:
        row = '{:<6} {:<35} {:<15} {:<15} {:<15}'
        logger.debug(row.format('Index', 'Name', 'Upper Limit', 'Lower Limit', 'Axis'))
        for index, name, up_limit, lo_limit, axis in data:
            logger.debug(row.format(index, name, up_limit, lo_limit, str(axis)))

    @property
    def observation_space(self) -> np.ndarray:
        return self._observation_space

    @property
    def action_space(self) -> np.ndarray:
        return self._action_space

    def close(self) -> None:
        self.sim.close()
