"""
CollisionObject / CollisionDetector — API-compatible with
/root/reference/robotic_manipulator_rloa/utils/collision_detector.py:9-61, answered by the batched
simulator's closest-distance kernel (rloa_sim_observe) instead of p.getClosestPoints.
`body` is the Environment's simulator, `obstacle_ids` are the tokens Environment hands out
('obstacle' / 'target').  Link-vs-link queries (compute_collisions_in_manipulator, :63-98) are answered by the
pair-GJK kernel (rloa_sim_self_distances).
"""
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class CollisionObject:
    body: object
    link: int


class CollisionDetector:

    def __init__(self, collision_object: CollisionObject, obstacle_ids: List[str]):
        self.obstacles = obstacle_ids
        self.collision_object = collision_object

    def compute_distances(self, max_distance: float = 10.0, env_index: int = 0) -> np.ndarray:
        """Closest distance from the link to each obstacle token; saturates at max_distance (:33-61)."""
        sim = self.collision_object.body
        _, link_obstacle, ee_target = sim.observe(want_distances=True)
        out = []
        for token in self.obstacles:
            if token == 'obstacle':
                d = float(link_obstacle[env_index, self.collision_object.link].item())
            elif token == 'target':
                if self.collision_object.link != sim.endeffector_index:
                    raise NotImplementedError('target distances are computed for the end-effector link only')
                d = float(ee_target[env_index].item())
            else:
                raise ValueError(f'unknown obstacle token {token!r}')
            out.append(min(d, max_distance))
        return np.array(out)

    def compute_collisions_in_manipulator(self, affected_joints: List[int], max_distance: float = 10.,
                                          env_index: int = 0) -> np.ndarray:
        """Closest distance from the link to each link in `affected_joints`, skipping the link itself and its two
        neighbours (always in contact), saturated at max_distance — also the value for a link without a collision
        shape, for which p.getClosestPoints returns nothing (:63-98)."""
        sim = self.collision_object.body
        link = self.collision_object.link
        row = sim.self_distances()[env_index, link].cpu().numpy().astype(float)
        out = []
        for joint_ind in affected_joints:
            if joint_ind in (link - 1, link, link + 1):
                continue
            out.append(min(float(row[joint_ind]), max_distance))
        return np.array(out)
