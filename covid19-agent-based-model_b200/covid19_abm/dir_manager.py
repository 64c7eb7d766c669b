"""Where the model's data files live.

Mirror of the reference's `covid19_abm.dir_manager.get_data_dir` (/root/reference/src/covid19_abm/dir_manager.py:3-8):
`<root>/data/<parts...>` where <root> is three directory levels above this file.  The census, CDR and
line-list inputs are private, so deployments usually point somewhere else: the environment variable
COVID19_ABM_DATA_DIR, when set, replaces `<root>/data`.
"""
import os

ENV_VAR = 'COVID19_ABM_DATA_DIR'


def get_data_dir(*args):
    base = os.environ.get(ENV_VAR)
    if not base:
        here = os.path.abspath(__file__)
        for _ in range(3):                      # covid19_abm/ -> package dir -> repository root
            here = os.path.dirname(here)
        base = os.path.join(here, 'data')
    return os.path.join(base, *args)
