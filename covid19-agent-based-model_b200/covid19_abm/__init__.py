"""Drop-in for the reference package `covid19_abm` (params, base_model, scenario_models, dir_manager) whose
daily agent step runs on a B200 through libabm_b200.so.  See DESIGN.md section 1 and INTEGRATION.md."""
__version__ = '0.1.0'
