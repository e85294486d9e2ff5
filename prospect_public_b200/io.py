"""Run-folder layout and file formats (prospect/io.py:11-127, prospect/profile.py:109-122).

<dir>/log.yaml, config.pkl, state.pkl, status.txt, analytical/random_gauss.pkl, profile/<p>.txt ...,
global_optimisation/results.txt, mcmc/<jobname>_<i>.txt + .paramnames + .ranges -- same names, same text
formats, so a folder written here can be read by PROSPECT's own `load_profile` and vice versa.
"""
import os
import pickle
import shutil

import numpy as np
import yaml


def read_user_input(arg):
    """yaml file -> (dict, False); run folder -> (its log.yaml with io.dir forced to the folder, True)."""
    if os.path.isfile(arg):
        with open(arg, 'r') as f:
            return yaml.full_load(f), False
    if os.path.isdir(arg):
        with open(os.path.join(arg, 'log.yaml'), 'r') as f:
            cfg = yaml.full_load(f)
        if cfg['io']['dir'] and arg != cfg['io']['dir']:
            cfg['io']['dir'] = arg
        return cfg, True
    raise ValueError('Invalid arguments to PROSPECT. Give either a .yaml input file or the folder of a previous '
                     'PROSPECT run.')


def prepare_run(config):
    """Create <dir>, <dir>/<jobtype>, <dir>/analytical and write log.yaml (io.py:49-71)."""
    if not config.io.write:
        return
    if os.path.isdir(config.io.dir) and config.io.overwrite_dir:
        shutil.rmtree(config.io.dir)
    os.makedirs(config.io.dir, exist_ok=False)
    os.makedirs(os.path.join(config.io.dir, config.run.jobtype))
    os.makedirs(os.path.join(config.io.dir, 'analytical'))
    plain = {}
    for sec, entries in config.config_dict.items():
        plain[sec] = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in entries.items()}
    with open(os.path.join(config.io.dir, 'log.yaml'), 'w') as f:
        yaml.dump(plain, f)


def dump_pickle(obj, folder, name):
    """Write through a backup file and rename (communication.py:159-166)."""
    tmp = os.path.join(folder, f'{name}_backup.pkl')
    with open(tmp, 'wb') as f:
        pickle.dump(obj, f)
    os.replace(tmp, os.path.join(folder, f'{name}.pkl'))


def load_pickle(folder, name):
    path = os.path.join(folder, f'{name}.pkl')
    if not os.path.isfile(path):
        raise ValueError(f'Could not find {path}.')
    with open(path, 'rb') as f:
        return pickle.load(f)


def load_config(folder):
    """The run's configuration: this package's pb200_config.pkl if present, else PROSPECT's own config.pkl
    (read through stand-in classes when PROSPECT is not installed)."""
    if os.path.isfile(os.path.join(folder, 'pb200_config.pkl')):
        return load_pickle(folder, 'pb200_config')
    from .compat import _Registered
    with _Registered():
        return load_pickle(folder, 'config')


def load_profile(prospect_folder, direct_txt=False):
    """{column name: array} of profile/<parameter>.txt (profile.py:109-122)."""
    if direct_txt:
        data_file = prospect_folder
    else:
        config = load_config(prospect_folder)
        data_file = f'{prospect_folder}/profile/{config.profile.parameter}.txt'
    if not os.path.isfile(data_file):
        raise KeyError(f"Could not load profile likelihood results: '{data_file}' not found.")
    results = np.genfromtxt(data_file, delimiter='\t', dtype=str, skip_header=0, autostrip=True)
    return {entry: np.array([float(val.strip('[]')) for val in results[1:, idx]])
            for idx, entry in enumerate(results[0])}


def write_parameters(param_dict, output_dir, jobname):
    """<jobname>.paramnames / .ranges with the leading blank line GetDist tolerates (io.py:92-122)."""
    names = [n.replace('*', '_') for n in param_dict]
    latex = []
    for n in names:
        us = [i for i, ch in enumerate(n) if ch == '_']
        latex.append(n[:us[0]] + '{' + n[us[0]:].replace('_', ',') + '}' if len(us) > 1 else n)
    with open(f'{output_dir}/{jobname}.paramnames', 'w') as f:
        for n, lx in zip(names, latex):
            f.write(f'\n{n} {lx}')
    with open(f'{output_dir}/{jobname}.ranges', 'w') as f:
        for n, item in zip(names, param_dict.values()):
            line = f'\n{n}'
            for bound in item['range']:
                line += ' N' if bound is None else f' {str(bound)}'
            f.write(line)


def write_chains(output_dir, jobname, *chain_list, first_index=0):
    for i, chain in enumerate(chain_list):
        np.savetxt(f'{output_dir}/{jobname}_{first_index + i}.txt', chain.data, delimiter=' ')


def unpack_mcmc(param_dict, output_dir, jobname, *chain_list, first_index=0, write_names=True):
    """`first_index` / `write_names`: a rank that owns chains [first_index, ...) of a sharded run writes only its own
    chain files; rank 0 also writes the .paramnames / .ranges pair."""
    os.makedirs(output_dir, exist_ok=True)
    if write_names:
        write_parameters(param_dict, output_dir, jobname)
    write_chains(output_dir, jobname, *chain_list, first_index=first_index)
