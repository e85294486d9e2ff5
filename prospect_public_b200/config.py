"""The yaml input schema (io / run / kernel / profile / mcmc sections), unchanged from PROSPECT so that
existing input files run as they are.

Mirrors `prospect.input.Configuration` (prospect/input.py:11-157) and the per-module `Arguments`
classes (prospect/io.py:133-172, run.py:135-169, kernels/base_kernel.py:144-195, profile.py:129-362,
mcmc.py:196-285): same keys, same defaults, same mandatory keys, same type / allowed-value checks,
unknown keys rejected, `profile.values` given as a string is evaluated with numpy only.
Here the schema is one declarative table instead of nested classes resolved by reflection.

Additions (the only ones): `run.mode` also accepts 'gpu' (the old values are accepted and all mean
"this process drives its GPU; launch with torchrun for several"), `kernel.type` also accepts
'cuda_gaussian' as a synonym of 'analytical'.
"""
import copy
import os
from types import SimpleNamespace

import numpy as np

_REQUIRED = object()
NoneType = type(None)
_LISTY = (list, np.ndarray)


class ConfigError(ValueError):
    pass


def safe_eval(code):
    """Evaluate a yaml string such as 'np.linspace(-1.0, 2.0, 10)' with numpy and nothing else
    (prospect/input.py:121-136)."""
    import builtins
    saved = builtins.__import__
    builtins.__import__ = None
    try:
        return eval(code, {'np': np, 'numpy': np})
    except Exception:
        raise ConfigError(f'Illegal input {code!r}. Only the standard and numpy library may be used in '
                          'Python-evaluated statements in the .yaml file.')
    finally:
        builtins.__import__ = saved


def _analytic(cfg):
    return cfg['kernel'].get('type') in ('analytical', 'cobaya', 'cuda_gaussian')


# (name, allowed python types or None, allowed values or None, default(cfg) or _REQUIRED)
_IO = [
    ('jobname', (str,), None, _REQUIRED),
    ('write', (bool,), None, _REQUIRED),
    ('dir', (str,), None, _REQUIRED),
    ('overwrite_dir', None, [True, False, 'yes', 'y'], lambda c: False),
    ('snapshot_interval', (float,), None, lambda c: 1.0),
]
_RUN = [
    ('jobtype', (str,), ['mcmc', 'profile', 'global_optimisation'], _REQUIRED),
    ('mode', (str,), ['mpi', 'threaded', 'serial', 'gpu'], _REQUIRED),
    ('num_processes', (int,), None, lambda c: 1),
    ('numpy_random_seed', (float, int, NoneType), None, lambda c: None),
]
_KERNEL = [
    ('type', (str,), ['analytical', 'montepython', 'cobaya', 'cuda_gaussian'], _REQUIRED),
    ('param', (str,), None, _REQUIRED),
    ('conf', (str,), None, lambda c: '' if _analytic(c) else _REQUIRED),
    ('path', (str,), None, lambda c: '' if _analytic(c) else _REQUIRED),
    ('debug', (bool,), None, lambda c: False),
    ('ignore_prior', (bool,), None, lambda c: False),
]


def _profile_parameter_default(c):
    if c['run']['jobtype'] == 'global_optimisation':
        return None
    raise ConfigError("You need to specify the 'parameter' input of the profile section.")


def _profile_values_default(c):
    if c['run']['jobtype'] == 'global_optimisation':
        return []
    raise ConfigError("You need to specify the 'values' input of the profile section.")


_PROFILE = [
    ('parameter', (str, NoneType), None, _profile_parameter_default),
    ('values', _LISTY, None, _profile_values_default),
    ('optimiser', None, ['simulated annealing'], lambda c: 'simulated annealing'),
    ('temperature_schedule', None, ['exponential'], lambda c: 'exponential'),
    ('temperature_range', _LISTY, None, _REQUIRED),
    ('step_size_schedule', None, ['exponential', 'adaptive'], lambda c: 'adaptive'),
    ('step_size_range', _LISTY + (NoneType,), None, lambda c: None),
    ('step_size_adaptive_interval', _LISTY + (NoneType,), None, lambda c: [0.05, 0.15]),
    ('step_size_adaptive_multiplier', (float, np.float64, str), None, lambda c: 'adaptive'),
    ('step_size_adaptive_initial', (float, np.float64), None, lambda c: 1.0),
    ('steps_per_iteration', (int, NoneType), None, lambda c: None),
    ('minutes_per_iteration', (int, float, NoneType), None, lambda c: None),
    ('max_iterations', (int, float), None, _REQUIRED),
    ('repetitions', (int, float), None, lambda c: 1),
    ('start_from_mcmc', (str, NoneType), None, lambda c: None),
    ('start_bin_fraction', (float,), None, lambda c: 0.1),
    ('chi2_tolerance', (float, int, NoneType), None, lambda c: None),
    ('plot_profile', (bool,), None, lambda c: True),
    ('detailed_plot', (bool,), None, lambda c: False),
    ('plot_Delta_chi2', (bool,), None, lambda c: True),
    ('plot_schedule', (bool,), None, lambda c: False),
    ('confidence_levels', (list,), None, lambda c: [0.68, 0.95]),
    ('start_from_covmat', (str, list, np.ndarray, NoneType), None, lambda c: None),
    ('start_from_position', (str, list, np.ndarray, NoneType), None, lambda c: None),
    ('start_from_profile', (str, NoneType), None, lambda c: None),
]
_MCMC = [
    ('algorithm', (str,), ['MetropolisHastings'], _REQUIRED),
    ('N_chains', (int,), None, _REQUIRED),
    ('steps_per_iteration', (int, NoneType), None, lambda c: None),
    ('minutes_per_iteration', (int, float, NoneType), None, lambda c: None),
    ('convergence_Rm1', (float,), None, _REQUIRED),
    ('unpack_at_dump', (bool,), None, _REQUIRED),
    ('temperature', (float,), None, lambda c: 1.0),
    ('step_size', (float,), None, lambda c: 1.0),
    ('analyse_automatically', None, [True, False, 'yes', 'y'], lambda c: False),
    ('start_from_covmat', (str, list, np.ndarray, NoneType), None, lambda c: None),
    ('start_from_position', (str, list, np.ndarray, NoneType), None, lambda c: None),
]
SCHEMA = {'io': _IO, 'run': _RUN, 'kernel': _KERNEL, 'profile': _PROFILE, 'mcmc': _MCMC}


def _validate_section(name, sec):
    """The cross-field `validate` hooks of the reference's InputArgument subclasses."""
    if name == 'kernel':
        if not os.path.isfile(os.path.join(os.getcwd(), sec['param'])):
            raise ConfigError(f"The file pointed to in the 'param' field of the 'kernel' input, which has the "
                              f"value {sec['param']}, could not be found.")
        # 'montepython' / 'cobaya' wrap external CPU codes: the yaml is valid (the reference's schema), the kernel object
        # has to be supplied by the caller -- run(yaml, kernel=hosted.HostKernelAdapter(...)); kernels.initialise_kernel
        # refuses to build one
        if sec['type'] == 'montepython' and not os.path.isdir(os.path.join(os.getcwd(), sec['path'])):
            raise ConfigError(f"The directory pointed to in the 'path' field of the 'kernel' input, which has the "
                              f"value {sec['path']}, could not be found.")          # base_kernel.py:176-180
    if name in ('profile', 'mcmc'):
        if sec['steps_per_iteration'] is not None and sec['minutes_per_iteration'] is not None:
            raise ConfigError("Give either 'steps_per_iteration' or 'minutes_per_iteration', not both.")
        if sec['minutes_per_iteration'] is not None and not sec['minutes_per_iteration'] > 0:
            raise ConfigError("'minutes_per_iteration' must be positive.")
    if name == 'profile':
        if sec['step_size_schedule'] == 'exponential' and sec['step_size_range'] is None:
            raise ConfigError("Input 'step_size_range' is set to None, but must be set when using the "
                              "'exponential' schedule type.")
        if isinstance(sec['step_size_adaptive_multiplier'], str) and \
                sec['step_size_adaptive_multiplier'] != 'adaptive':
            raise ConfigError("The only allowed string input for 'step_size_adaptive_multiplier' is 'adaptive'.")
        if sec['step_size_schedule'] == 'adaptive':
            sec['step_size_range'] = [sec['step_size_adaptive_initial'], None]      # profile.py:210-213
        if sec['start_from_profile'] is not None and sec['start_from_mcmc'] is None and \
                sec['start_from_covmat'] is None:
            raise ConfigError("'start_from_profile' needs 'start_from_mcmc' or 'start_from_covmat' for a covmat.")
    if name == 'mcmc':
        if sec['analyse_automatically'] and not sec['unpack_at_dump']:
            raise ConfigError("Must unpack MCMC in order to analyse; please set 'unpack_at_dump' to True.")


class Configuration:
    """config.io / .run / .kernel / .profile | .mcmc namespaces + config_dict (what log.yaml stores)."""

    def __init__(self, config_yaml, quiet=True):
        cfg = copy.deepcopy(config_yaml)
        for sec in ('io', 'run', 'kernel'):
            cfg.setdefault(sec, {})
        jobtype = cfg['run'].get('jobtype')
        if jobtype is None:
            raise ConfigError("You must give 'jobtype' as an input under the 'run' category.")
        if jobtype == 'mcmc':
            sections = ['run', 'io', 'kernel', 'mcmc']
        elif jobtype in ('profile', 'global_optimisation'):
            sections = ['run', 'io', 'kernel', 'profile']
        else:
            raise ConfigError("The given value of 'jobtype' is not recognised. Choose either 'mcmc', 'profile' "
                              "or 'global_optimisation'.")
        for sec in sections:
            cfg.setdefault(sec, {})
        if isinstance(cfg.get('profile', {}).get('values'), str):
            cfg['profile']['values'] = safe_eval(cfg['profile']['values'])
        for sec in sections:
            known = {row[0] for row in SCHEMA[sec]}
            extra = set(cfg[sec]) - known
            if extra:
                raise ConfigError(f"Unknown input(s) {sorted(extra)} in section '{sec}'.")
        # defaults may depend on other sections: iterate to a fixed point (input.py:54-79)
        missing = []
        for _ in range(4):
            missing = []
            for sec in sections:
                for name, _types, _allowed, default in SCHEMA[sec]:
                    if name in cfg[sec]:
                        continue
                    if default is _REQUIRED:
                        missing.append(f'{sec}.{name}')
                        continue
                    try:
                        val = default(cfg)
                    except KeyError:
                        missing.append(f'{sec}.{name}')
                        continue
                    if val is _REQUIRED:
                        missing.append(f'{sec}.{name}')
                    else:
                        cfg[sec][name] = val
            if not missing:
                break
        if missing:
            raise ConfigError(f'Could not set default value for inputs {missing}: they have no default and must '
                              'be specified.')
        for sec in sections:
            for name, types, allowed, _default in SCHEMA[sec]:
                val = cfg[sec][name]
                if types is not None and type(val) not in types:
                    raise ConfigError(f"Input '{name}' was given type value of {type(val)} which is not one of "
                                      f"the allowed types: {types}.")
                if allowed is not None and val not in allowed:
                    raise ConfigError(f"Input '{name}' was given value {val} which not in the list of allowed "
                                      f"values: {allowed}")
            _validate_section(sec, cfg[sec])
        self._set_output_dir(cfg['io'])
        self.config_dict = cfg
        for sec in sections:
            setattr(self, sec, SimpleNamespace(**cfg[sec]))

    @staticmethod
    def _set_output_dir(io):
        """Never clobber an existing run folder unless asked to (input.py:108-119)."""
        if io['write'] and not io['overwrite_dir'] and os.path.isdir(io['dir']):
            idx = 0
            while os.path.isdir(f"{io['dir']}_{idx}"):
                idx += 1
            io['dir'] = f"{io['dir']}_{idx}"

    @property
    def seed(self):
        """run.numpy_random_seed doubles as the Philox key; None draws one from the OS -- ONCE per Configuration, so
        that every use (the optimiser, the snapshot a resumed run replays from) sees the same key.  Under torchrun
        `synchronise()` makes rank 0's value the job's."""
        s = self.run.numpy_random_seed
        if s is not None:
            return int(s) & 0xFFFFFFFFFFFFFFFF
        if getattr(self, '_drawn_seed', None) is None:
            self._drawn_seed = int.from_bytes(os.urandom(8), 'little')
        return self._drawn_seed

    def synchronise(self):
        """Make rank 0's resolved output folder and Philox key those of every rank of a torch.distributed job (every
        rank builds its own Configuration: a late rank could otherwise de-duplicate io.dir differently, and ranks
        without a numpy_random_seed would each draw their own key).  No-op for a single process."""
        from . import distributed
        rank, world, active = distributed.dist_info()
        if not active or world == 1:
            return self
        import torch.distributed as dist
        payload = [{'dir': self.io.dir, 'seed': self.seed}]
        dist.broadcast_object_list(payload, src=0)
        self.io.dir = self.config_dict['io']['dir'] = payload[0]['dir']
        if self.run.numpy_random_seed is None:
            self._drawn_seed = int(payload[0]['seed'])
        return self
