"""Consumer side of the daily log (reference: src/covid19_abm/data_plotting.py).

`load_model_logs_df` is what the reference's notebooks call on a `model_log_file_*.log` written by
`Country.log_model_output` (base_model.py:1082-1106); it is the only downstream reader of the hot path's
output, so the drop-in keeps it.  The plotting helpers of the reference module need
matplotlib / seaborn / geopandas; they are presentation code outside the hot path: the panel list is kept
(`set_plot_config`), the drawing function says so when called.
"""
import json

import numpy as np
import pandas as pd

__all__ = ['load_model_logs_df', 'load_scenario_logs_df', 'set_plot_config', 'plot_trajectories']


def _read_rows(log_name, delim='$$'):
    sep = f' {delim} '
    with open(log_name) as fl:
        text = fl.read().strip()
    return [json.loads(line.split(sep)[-1]) for line in text.split('\n')] if text else []


def load_model_logs_df(log_name):
    """One row per logged day, in file order: the 12 counters + `date` of the log line, then `new_cases`,
    `new_hospitalized`, `new_critical`, `version` (= log_name) and `days` (0-based row number).

    data_plotting.py:30-64.  Kept as the reference computes them: `new_cases` and `new_hospitalized` are first
    differences of the cumulative counts (first row = the count itself); `new_critical` is
    `critical_count - previous row's infected_count` (the reference carries `infected_count` into its
    `prev_critical`, data_plotting.py:55) -- downstream notebooks were written against that column as is.
    """
    rows = _read_rows(log_name)
    df = pd.DataFrame(rows)
    if not rows:
        return df
    inf = df['infected_count'].to_numpy(np.int64)
    hosp = df['hospitalized_count'].to_numpy(np.int64)
    crit = df['critical_count'].to_numpy(np.int64)
    prev_inf = np.concatenate(([0], inf[:-1]))
    df['new_cases'] = inf - prev_inf
    df['new_hospitalized'] = hosp - np.concatenate(([0], hosp[:-1]))
    df['new_critical'] = crit - prev_inf
    df['version'] = log_name
    df['days'] = np.arange(len(df))
    return df


def load_scenario_logs_df(log_names, scenario=None, scale=1):
    """Convenience on top of `load_model_logs_df` (no reference counterpart: the notebooks do this inline,
    notebooks/20200630 cell 9): concatenate the ensemble members' logs and add the columns the notebooks plot
    (`Deaths`, `Infected`, `New cases`, `Active cases` = currently contagious, `Hospitalizations`,
    `Current hospitalizations`, `ICU admissions`, `Current ICU admissions`), each `scale` x the logged counter
    (scale = 100 / sample percentage), plus `scenario`."""
    frames = [load_model_logs_df(n) for n in log_names]
    df = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame()
    if len(df):
        df['date'] = pd.to_datetime(df['date'])
        for column, counter in (('Deaths', 'died_count'), ('Infected', 'infected_count'), ('New cases', 'new_cases'),
                                ('Active cases', 'current_contagious_cases'), ('Hospitalizations', 'hospitalized_count'),
                                ('Current hospitalizations', 'current_hospitalized_cases'),
                                ('ICU admissions', 'critical_count'), ('Current ICU admissions', 'current_critical_cases')):
            df[column] = scale * df[counter]
        if scenario is not None:
            df['scenario'] = scenario
    return df


def set_plot_config(df):
    """The eight panels of the reference's scenario figures (data_plotting.py:67-79) as keyword sets for
    `plot_trajectories`."""
    panels = [('Deaths', 'Cumulative deaths', 'Number of deaths', 'upper left'),
              ('Infected', 'Cumulative infected incidence', 'Number of infected', 'upper left'),
              ('New cases', 'Daily new cases', 'New cases', 'upper right'),
              ('Active cases', 'Daily active cases', 'Active cases', 'upper right'),
              ('Hospitalizations', 'Cumulative hospitalizations', 'Number of hospitalized', 'upper left'),
              ('Current hospitalizations', 'Current hospitalizations', 'Number of hospitalized', 'upper right'),
              ('ICU admissions', 'Cumulative ICU admissions', 'Number of admissions', 'upper left'),
              ('Current ICU admissions', 'Current ICU admissions', 'Number of admissions', 'upper right')]
    return [dict(x='date', y=y, hue='scenario', data=df, title=title, y_label=y_label, legend_loc=loc)
            for y, title, y_label, loc in panels]


def plot_trajectories(*args, **kwargs):
    """Presentation code (data_plotting.py:82-101): needs matplotlib + seaborn, which are not part of the hot path
    and are absent from the build image."""
    raise ImportError('covid19_abm.data_plotting.plot_trajectories needs matplotlib and seaborn; the B200 drop-in '
                      'ships the log reader only (load_model_logs_df / set_plot_config)')
