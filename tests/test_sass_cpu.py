"""
Static check (no GPU): the built library carries sm_100a code whose SASS shows what DESIGN.md claims — tcgen05 MMAs
(UTC*MMA), 1-D TMA bulk copies (UBLKCP), TMEM loads (LDTM), mbarrier waits (SYNCS) in the tensor-core kernels and packed
FFMA2 in the Gauss-Seidel sweep of the simulator.  Mnemonics per /opt/skills/guides/B200_PROFILING.md.
"""
import shutil
import subprocess

import pytest

from robotic_manipulator_rloa_b200 import _native

CUOBJDUMP = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'


def sass(pattern):
    try:
        listing = subprocess.check_output([CUOBJDUMP, '-lelf', _native.LIB_PATH], text=True)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip('cuobjdump not available')
    assert 'sm_100a' in listing
    out = subprocess.check_output([CUOBJDUMP, '-sass', '-fun', pattern, _native.LIB_PATH], text=True,
                                  stderr=subprocess.STDOUT)
    assert 'Function' in out, f'{pattern} not found in the library'
    return out


def count(text, mnemonic):
    return sum(1 for line in text.splitlines() if mnemonic in line)


def test_policy_kernel_is_tcgen05_with_bulk_copies():
    s = sass('_ZN4rloa20policy_act_tc_kernelEPKhPKfiiiyyPKyfPf')
    assert count(s, 'UTCHMMA') == 6 + 16 + 16          # tf32 layer 1 (3 k-steps x hi / lo), bf16 layer 2, bf16 heads
    assert count(s, 'UBLKCP') == 4                     # W1 image, vectors, W2 image, head image
    assert count(s, 'LDTM') >= 2 and count(s, 'SYNCS.PHASECHK') >= 4 and count(s, 'UTCBAR') >= 3


@pytest.mark.parametrize('mode,mmas', [(0, 16), (1, 16), (2, 6)])
def test_trunk_kernels_are_tcgen05(mode, mmas):
    s = sass(f'_ZN4rloa22trunk_tc_layer2_kernelILi{mode}EEEvNS_7TcBatchEii')
    assert count(s, 'UTCHMMA') == mmas and count(s, 'LDTM') >= 1 and count(s, 'UTCBAR') >= 1


def test_solve_kernel_uses_packed_ffma2():
    s = sass('_ZN4rloa16sim_solve_kernelILi12ELi16ELb1ELb1ELb0EEEvNS_8ModelDevENS_9SimArraysENS_10StepCfgDevEiPKfPKhPfS8_PhS9_')
    assert count(s, 'FFMA2') >= 288                    # 12 rows x 6 pairs x (2 sweep directions + limit rows)
    assert count(s, 'UTCHMMA') == 0                    # nothing GEMM-shaped in the simulator, by design


def test_learn_cluster_kernel_is_tcgen05_with_cluster_barriers_and_system_scope_words():
    """NAFAgent.learn as one kernel: every contraction of forward and backward is a tcgen05 MMA (layer 1 twice as tf32 hi / lo,
    layer 2, heads, dWh^T, da2, da1, dW1, dW2), operands arrive by TMA bulk copies, the BatchNorm statistics cross the cluster
    (barrier.cluster arrive / wait pairs), and the in-kernel gradient exchange moves system-scope 64-bit words."""
    s = sass('_ZN4rloa24naf_learn_cluster_kernelENS_16LearnClusterArgsE')
    assert count(s, 'UTCHMMA') >= 100
    assert count(s, 'UBLKCP') >= 5 and count(s, 'LDTM') >= 10 and count(s, 'UTCBAR') >= 8
    assert count(s, 'UCGABAR_ARV') >= 5 and count(s, 'UCGABAR_WAIT') >= 5
    assert count(s, 'STRONG.SYS') >= 20


def test_collision_phase_and_deferred_replay_commit_kernels_exist():
    sass('_ZN4rloa19sim_contacts_kernelILi16EEEvNS_8ModelDevENS_9SimArraysEf')
    sass('_ZN4rloa20replay_commit_kernelE11rloa_replayiPKh')
