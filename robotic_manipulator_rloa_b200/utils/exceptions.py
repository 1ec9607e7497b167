"""
Exception types of the framework — same class names, hierarchy and default messages as the reference
(/root/reference/robotic_manipulator_rloa/utils/exceptions.py:6-135) so user `except` clauses and
message checks stay drop-in.  Built from one small factory instead of eight hand-written classes.
"""
from __future__ import annotations

from typing import Optional


class FrameworkException(Exception):
    """Base class of every framework error; ``str(e)`` is ``ClassName: message``."""

    def __init__(self, message: str) -> None:
        Exception.__init__(self, message)
        self.message = message

    def __str__(self) -> str:
        return f'{self.__class__.__name__}: {self.message}'

    def set_message(self, value: str) -> 'FrameworkException':
        self.message = value
        return self


def _framework_error(name: str, default_message: str, doc: str):
    def __init__(self, message: Optional[str] = None) -> None:
        if message:
            self.message = message
        FrameworkException.__init__(self, self.message)

    return type(name, (FrameworkException,), {'message': default_message, '__init__': __init__, '__doc__': doc,
                                              '__module__': __name__})


InvalidManipulatorFile = _framework_error(
    'InvalidManipulatorFile', 'The URDF/SDF file received is not valid',
    'The URDF/SDF file cannot be turned into a simulator model.')
InvalidHyperParameter = _framework_error(
    'InvalidHyperParameter', 'The hyperparameter received is not valid',
    'set_hyperparameter() got an unknown name or an out-of-range value.')
InvalidEnvironmentParameter = _framework_error(
    'InvalidEnvironmentParameter', 'The Environment parameter received is not valid',
    'The Environment was configured with an invalid parameter.')
InvalidNAFAgentParameter = _framework_error(
    'InvalidNAFAgentParameter', 'The NAF Agent parameter received is not valid',
    'The NAFAgent was configured with an invalid parameter.')
EnvironmentNotInitialized = _framework_error(
    'EnvironmentNotInitialized',
    'The Environment is not yet initialized. The environment can be initialized via the '
    'initialize_environment() method',
    'A method that needs the Environment was called before initialize_environment().')
NAFAgentNotInitialized = _framework_error(
    'NAFAgentNotInitialized',
    'The NAF Agent is not yet initialized. The agent can be initialized via the initialize_naf_agent() method',
    'A method that needs the NAFAgent was called before initialize_naf_agent().')
MissingWeightsFile = _framework_error(
    'MissingWeightsFile', 'The weight file provided does not exist',
    'Pretrained weights were requested from a path that does not exist.')
ConfigurationIncomplete = _framework_error(
    'ConfigurationIncomplete',
    'The configuration for the training is incomplete. Either the Environment, the '
    'NAF Agent or both are not yet initialized. The environment can be initialized via the '
    'initialize_environment() method, and the agent can be initialized via the '
    'initialize_naf_agent() method',
    'run_training() / test_trained_model() was called without an Environment and a NAFAgent.')
