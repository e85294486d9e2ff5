#!/usr/bin/env python
"""A profile likelihood for a likelihood that only exists on the HOST, through the batched GPU machinery.

The kernel below has PROSPECT's kernel plugin surface (prospect/kernels/base_kernel.py:9-138: `param`, `loglkl(position
dict)`, `logprior`, `set_fixed_parameters`, `read_initial_position`, `read_covmat`, defaults) and a non-Gaussian
"banana" likelihood written in plain numpy -- a stand-in for a cobaya / MontePython kernel.  `HostKernelAdapter` turns
it into a batched likelihood; proposals, prior-box tests, accept decisions, running bestfits and the annealing schedule
of all (value x repetition) chains run on the GPU, and the usual PROSPECT output tree is
written.

    python examples/host_kernel_example.py [output_dir]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prospect_public_b200.hosted import HostKernelAdapter  # noqa: E402
from prospect_public_b200.io import load_profile  # noqa: E402
from prospect_public_b200.run import run  # noqa: E402


class BananaKernel:
    """-log L = 1/2 [ (a - 1)^2 / 0.5^2 + (b - a^2)^2 / 0.2^2 + sum_i c_i^2 / 0.3^2 ],  parameters a, b, c1..c3."""

    def __init__(self):
        self.names = ['a', 'b', 'c1', 'c2', 'c3']
        self.param = {'varying': {nm: {'range': [-4.0, 4.0]} for nm in self.names}, 'fixed': {}, 'derived': {}}

    def set_fixed_parameters(self, fixed_param_dict):                      # base_kernel.py:46-51
        for name, value in fixed_param_dict.items():
            self.param['fixed'][name] = self.param['varying'].pop(name)
            self.param['fixed'][name]['fixed_value'] = value

    @property
    def varying_param_names(self):
        return list(self.param['varying'].keys())

    def loglkl(self, prop):                                                # base_kernel.py:72-81
        for name, par in self.param['fixed'].items():
            prop[name] = [par['fixed_value']]
        a, b = prop['a'][0], prop['b'][0]
        c = np.array([prop[f'c{i}'][0] for i in (1, 2, 3)])
        return 0.5 * (((a - 1.0) / 0.5) ** 2 + ((b - a * a) / 0.2) ** 2 + np.sum((c / 0.3) ** 2))

    def logprior(self, position):                                          # base_kernel.py:83-102
        for name, val in position.items():
            if name in self.param['varying']:
                lo, hi = self.param['varying'][name]['range']
                if val[0] < lo or val[0] > hi:
                    return np.inf
        return 0.0

    def read_initial_position(self, config_initial_position):
        return np.array(config_initial_position, dtype=float)

    def read_covmat(self, config_covmat):
        return np.array(config_covmat, dtype=float)

    def finalize(self):
        pass


def analytic_profile(a):
    """min over (b, c) of -log L at fixed a: b = a^2, c = 0."""
    return 0.5 * ((a - 1.0) / 0.5) ** 2


def main(out_dir):
    values = [0.0, 0.5, 1.0, 1.5, 2.0]
    job = {
        'io': {'jobname': 'banana', 'write': True, 'dir': out_dir, 'overwrite_dir': True, 'snapshot_interval': 1.0e9},
        'run': {'jobtype': 'profile', 'mode': 'gpu', 'numpy_random_seed': 3},
        # the yaml's kernel section is only validated (the `param` file must exist); the kernel object is handed over below
        'kernel': {'type': 'cobaya', 'param': os.path.abspath(__file__)},
        'profile': {
            'parameter': 'a', 'values': values, 'optimiser': 'simulated annealing', 'repetitions': 4,
            'temperature_schedule': 'exponential', 'temperature_range': [0.5, 1.0e-4], 'max_iterations': 25,
            'steps_per_iteration': 100, 'step_size_schedule': 'adaptive', 'step_size_adaptive_interval': [0.19, 0.21],
            'step_size_adaptive_multiplier': 0.3, 'step_size_adaptive_initial': 1.0,
            'start_from_position': [1.0, 1.0, 0.0, 0.0, 0.0],
            'start_from_covmat': np.diag([0.25, 0.04, 0.09, 0.09, 0.09]).tolist(),
            'plot_profile': False, 'plot_schedule': False,
        },
    }
    res = run(job, kernel=lambda config_kernel, output_folder, rank, device: HostKernelAdapter(BananaKernel(), device=device))
    prof = load_profile(out_dir)
    print(f"{res['chain_steps']} chain-steps, {res['optimiser'].kernel.evaluations} host likelihood calls, "
          f"{res['time_to_result_s']:.2f} s; profile written to {os.path.join(out_dir, 'profile', 'a.txt')}")
    print('    a      -logL (annealed)   -logL (analytic)        b (a^2)')
    worst = 0.0
    for a, ll, b in zip(prof['a'], prof['-loglkl'], prof['b']):
        print(f'{a:6.2f}   {ll:16.6f}   {analytic_profile(a):16.6f}   {b:8.4f} ({a * a:.4f})')
        worst = max(worst, ll - analytic_profile(a))
    assert -1e-9 < worst < 5e-2, worst           # T_final = 1e-4: the chains end within a few T of the minimum
    return res


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join('output', 'host_kernel_example'))
