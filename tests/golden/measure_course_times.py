"""Disease-course times of 200 000 infections drawn by the REAL reference (`Country.set_epidemic_status`, numpy gamma /
exponential draws, base_model.py:838-1055) against the oracle / kernel form (Philox words through the host-built quantile
tables).  Build container only (needs /root/reference):  python tests/golden/measure_course_times.py
"""
import sys, os
sys.path[:0] = ['/root/repo/tests/golden', '/root/repo']
import numpy as np
from datetime import datetime
import ref_harness as rh
from abm_b200 import synth, tables
from oracle.abm_oracle import OracleModel
start = datetime(2020, 9, 1, 0)
spec0 = synth.make_spec(n_districts=10, R0=1.9, seed=5, start_datetime=start)
pop = synth.make_population(200_000, spec0, seed=5, scenario_flags=True)
w = rh.RefWorld(spec0, pop, seed_cases={0: 1})
np.random.seed(3)
m = w.make_model('UnmitigatedScenario', seed_infected=1, infect=False)
spec = rh.spec_from_reference_params(m.params, 240, start)
ids = np.arange(pop.n)
m.set_epidemic_status(ids)                                  # native numpy RNG
from datetime import timedelta
def ticks(a):
    out = np.full(len(a), 2**31 - 1, dtype=np.float64)
    for i, d in enumerate(a):
        if d != datetime.max:
            out[i] = (d - start) / timedelta(minutes=1)
    return out
ref = {k: ticks(getattr(m, k)) for k in ('date_start_contagious', 'date_start_symptomatic', 'date_recovered', 'date_start_hospitalized', 'date_start_critical', 'date_died')}
ref['infected_symptomatic_status'] = np.asarray(m.infected_symptomatic_status)
o = OracleModel(pop, dict(od_cum=spec.od_cum, mild_prob=1.0), tables.build_tables(spec), 99)
o.set_epidemic_status(ids)
ora = o.state_dict()
T = 2**31 - 1
def stats(name, a, b):
    a, b = a[a < T].astype(float), b[b < T].astype(float)
    print(f'{name:28s} ref n={a.size:7d} mean {a.mean()/1440:8.4f} d sd {a.std()/1440:7.4f}   oracle n={b.size:7d} mean {b.mean()/1440:8.4f} d sd {b.std()/1440:7.4f}   diff {100*(b.mean()/a.mean()-1):+.2f} %')
for k in ('date_start_contagious', 'date_start_symptomatic', 'date_recovered', 'date_start_hospitalized', 'date_start_critical', 'date_died'):
    stats(k, ref[k], ora[k])
for k in ('date_recovered',):
    a = (ref[k].astype(float) - ref['date_start_contagious'])[ref[k] < T]; b = (ora[k].astype(float) - ora['date_start_contagious'])[ora[k] < T]
    print('contagious period (recovered - contagious): ref %.4f d  oracle %.4f d  diff %+.2f %%' % (a.mean()/1440, b.mean()/1440, 100*(b.mean()/a.mean()-1)))
sym_r, sym_o = (ref['infected_symptomatic_status'] == 1), (ora['infected_symptomatic_status'] == 1)
print('symptomatic share ref %.4f oracle %.4f' % (sym_r.mean(), sym_o.mean()))
for nm, sel_r, sel_o in (('sym', sym_r, sym_o), ('asym', ~sym_r, ~sym_o)):
    a = (ref['date_recovered'].astype(float) - ref['date_start_contagious'])[sel_r & (ref['date_recovered'] < T)]
    b = (ora['date_recovered'].astype(float) - ora['date_start_contagious'])[sel_o & (ora['date_recovered'] < T)]
    print(f'  contagious period {nm}: ref {a.mean()/1440:.4f} d  oracle {b.mean()/1440:.4f} d  diff {100*(b.mean()/a.mean()-1):+.2f} %')
    a = ref['date_start_contagious'].astype(float)[sel_r]; b = ora['date_start_contagious'].astype(float)[sel_o]
    print(f'  latent period {nm}:     ref {a.mean()/1440:.4f} d  oracle {b.mean()/1440:.4f} d  diff {100*(b.mean()/a.mean()-1):+.2f} %')
w.close()
