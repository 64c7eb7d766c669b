"""State codes, bit layout and epidemiological constants of the agent step.

Values restate the reference's configuration (file:line are relative to /root/reference):
  * state codes                      src/covid19_abm/base_model.py:528-543, 629-630
  * age-band tables                  src/covid19_abm/params.py:54-122
  * dwell distributions, periods     src/covid19_abm/params.py:145-189
  * movement probabilities           src/covid19_abm/params.py:13-38
They are numbers, not code; the derivations (band -> per-age lookup, rate formula) are
re-implemented in `tables.py`.
"""
import numpy as np

# ----------------------------------------------------------------------------- state codes
STATE_SUSCEPTIBLE, STATE_INFECTED, STATE_CONTAGIOUS, STATE_RECOVERED, STATE_DEAD = 0, 1, 2, 3, 4
CLINICAL_NOT_HOSPITALIZED, CLINICAL_HOSPITALIZED, CLINICAL_CRITICAL, CLINICAL_RELEASED_OR_DEAD = 0, 1, 2, 3
SYMPTOM_NONE, SYMPTOM_ASYMPTOMATIC, SYMPTOM_SYMPTOMATIC = -1, 0, 1
DEAD_LOCATION_ID = -1
DISTRICT_MOVER_FALSE, DISTRICT_MOVER_TRUE = 0, 1

# ----------------------------------------------------------------------------- time base
TICKS_PER_HOUR = 60            # 1 tick = 1 minute
TICKS_PER_DAY = 24 * 60
T_NEVER = np.int32(2**31 - 1)  # datetime.max in the reference
FIRE_NEVER = 0xFFFF            # "no pending event" in the 16-bit next-event field

# ----------------------------------------------------------------------------- hot-record bit layout
# flags word (uint32) -- one per agent slot, read every sub-step.  Fields the kernel tests together sit
# next to each other (state + location in the low byte; byte 2 is the index into the rate tables).
F_EPI_SHIFT, F_EPI_MASK = 0, 0x7          # epidemic state 0..4
F_CLIN_SHIFT, F_CLIN_MASK = 3, 0x3        # clinical state 0..3
F_LOC_SHIFT, F_LOC_MASK = 5, 0x7          # location kind (LOC_*)
F_AWAY = 1 << 8                           # return_district_at != max
F_DMOVER = 1 << 9                         # district_mover
F_HEAD = 1 << 10                          # first member of its household (segment head)
F_ONSET = 1 << 11                         # symptom onset has passed (date_start_symptomatic < now)
F_STATUS_SHIFT, F_STATUS_MASK = 12, 0xF   # economic status id
F_AGE_SHIFT, F_AGE_MASK = 16, 0x7F        # age 0..99
F_SYMPT = 1 << 23                         # infected_symptomatic_status == SYMPTOMATIC  (age | F_SYMPT = rate index)
F_ASYMPT = 1 << 24                        # infected_symptomatic_status == ASYMPTOMATIC
F_OUT_H = 1 << 25                         # will be hospitalised
F_OUT_C = 1 << 26                         # will need critical care
F_OUT_D = 1 << 27                         # will die
F_EKIND_SHIFT, F_EKIND_MASK = 28, 0x3     # economic-activity location kind (EKIND_*)
F_BIGHH = 1 << 30                         # member of a household larger than BIG_HOUSEHOLD
F_FILLER = 1 << 31                        # padding slot, not a person
# aux word (uint32): high 16 = sub-step index at which the next scheduled event fires (FIRE_NEVER = none),
#                    low 16  = current district id
AUX_FIRE_SHIFT = 16

# location kinds say WHERE the agent is (the reference's location id follows from it):
#   HOME      own household                                   -> household id
#   HOMEPOOL  outside pool of its HOME district               -> home district id (economic activity there,
#             also for travellers that are away on paper: quirk Q4; base_model.py:331-339)
#   DEST      outside pool of the destination of a trip       -> current_district id (base_model.py:436-439)
#   POOL      extra outside location (school)                 -> its economic_activity_location id
#   HOSP      hospital = -current_district (base_model.py:270-272);  DEAD = -1
LOC_HOME, LOC_HOMEPOOL, LOC_DEST, LOC_POOL, LOC_HOSP, LOC_DEAD = 0, 1, 2, 3, 4, 5
EKIND_HOUSEHOLD, EKIND_HOME_DISTRICT, EKIND_POOL = 0, 1, 2

TILE = 1024           # agent slots per thread block (8 warps x SUBTILE)
SUBTILE = 128         # slots per warp: whole households (<= BIG_HOUSEHOLD members) of ONE home district; slot
                      # offset o of a sub-tile holds canonical agent id base + o with base a multiple of 4
BIG_HOUSEHOLD = 32    # larger households (mining camps, scenario_models.py:232-236) use the atomic path
MAX_STATUS = 12       # packed headcount reduction holds 12 statuses (3 words x 4 x 8 bit)

# ----------------------------------------------------------------------------- epidemiology (params.py)
SYMPTOMATIC_RATE = 0.6                      # params.py:179
CRITICAL_FATALITY_RATE = 0.5                # params.py:182
ASYM_CONTAGIOUS_MEAN_DAYS = 6.5             # params.py:145
SYM_CONTAGIOUS_MEAN_DAYS = 7.0              # params.py:149
ASYM_RECOVERY_GAMMA = (8.45, 0.7692307692307693)        # params.py:146-147 (shape, scale in days)
SYM_RECOVERY_GAMMA = (9.8, 0.7142857142857143)          # params.py:150-151
INCUBATION_GAMMA = (6.474197530864197, 0.7074235807860262)  # params.py:188-189
ASYM_TO_CONTAGIOUS_MEAN_DAYS = 4.6          # params.py:154
ONSET_TO_CONTAGIOUS_TICKS = -12 * 60        # base_model.py:881
ONSET_TO_HOSPITAL_TICKS = 5 * TICKS_PER_DAY     # params.py:165
HOSPITAL_TICKS = 8 * TICKS_PER_DAY              # params.py:169
CRITICAL_TICKS = 8 * TICKS_PER_DAY              # params.py:173
ONSET_TO_DEATH_TICKS = 21 * TICKS_PER_DAY       # params.py:177
HOUSEHOLD_RATE_FACTOR = 3.0                 # base_model.py:184-187
DISTRICT_MOVEMENT_ALLOWED_AGE = 18          # params.py:197
GO_OUT_HOUR, GO_HOME_HOUR, LOG_HOUR = 8, 16, 8   # params.py:131-136, base_model.py:1083

# (upper age bound inclusive, value); the last band is open-ended.  params.py:54-122
HOSPITALIZATION_PCT_BY_AGE = [(14, 0.1), (19, 0.2), (24, 0.5), (29, 1.0), (34, 1.6), (39, 2.3), (44, 2.9),
                              (49, 3.9), (54, 5.8), (59, 7.2), (64, 10.2), (69, 11.7), (74, 14.6),
                              (79, 17.7), (np.inf, 18.0)]
CRITICAL_PCT_BY_AGE = [(34, 5.0), (39, 5.3), (44, 6.0), (49, 7.5), (54, 10.4), (59, 14.9), (64, 22.4),
                       (69, 30.7), (74, 38.6), (79, 46.1), (np.inf, 70.9)]
HOSPITAL_DEATH_PCT_BY_AGE = [(39, 1.3), (44, 1.5), (49, 1.9), (54, 2.7), (59, 4.2), (64, 6.9), (69, 10.5),
                             (74, 14.9), (79, 20.3), (np.inf, 58.0)]
PER_CAPITA_CONTACT_RATE_BY_AGE = [(5, 7.4709), (10, 10.4079), (15, 12.4579), (20, 11.1542), (25, 10.8716),
                                  (30, 9.7019), (35, 11.7569), (40, 12.5434), (45, 13.2094), (50, 10.7766),
                                  (55, 10.3643), (60, 10.9087), (65, 11.1501), (70, 11.6564), (75, 11.8107),
                                  (np.inf, 11.5379)]

# Economic statuses in the order of the shipped interaction matrix
# (data/raw/final_close_interaction_matrix.xlsx).
ECONOMIC_STATUS_NAMES = [
    'Not working, inactive, not in universe', 'In School', 'Homemakers/Housework', 'Office workers',
    'Service Workers', 'Agriculture Workers', 'Indusrtry Workers', 'In the army', 'Disabled and not working']
WEEKDAY_MOVE_PROB = {  # params.py:13-24
    'Not working, inactive, not in universe': 0.8, 'In School': 1.0, 'Homemakers/Housework': 0.8,
    'Office workers': 1.0, 'Teachers': 1.0, 'Service Workers': 1.0, 'Agriculture Workers': 1.0,
    'Indusrtry Workers': 1.0, 'In the army': 1.0, 'Disabled and not working': 0.0}
OTHER_DAY_MOVE_PROB = {  # params.py:27-38
    'Not working, inactive, not in universe': 0.8, 'In School': 0.8, 'Homemakers/Housework': 0.8,
    'Office workers': 0.8, 'Teachers': 0.8, 'Service Workers': 0.8, 'Agriculture Workers': 0.8,
    'Indusrtry Workers': 0.8, 'In the army': 0.8, 'Disabled and not working': 0.0}
INTERACTION_SIZES = [15, 30, 15, 20, 50, 10, 25, 25, 5]
INTERACTION_MATRIX = [
    [.8, 0, .04, .04, .04, .04, .04, 0, 0],
    [0, .9, 0, .1, 0, 0, 0, 0, 0],
    [.05, 0, .7, .05, .2, 0, 0, 0, 0],
    [.01, .03, .03, .8, .1, 0, .02, .01, 0],
    [.05, 0, .2, .1, .6, .01, .03, .01, 0],
    [.02, 0, 0, 0, .08, .9, 0, 0, 0],
    [.02, 0, 0, .08, .05, 0, .85, 0, 0],
    [0, 0, 0, .02, .03, 0, 0, .95, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 1]]

# 2012 census totals of the 60 "new" districts d_1..d_60 (data/raw/district_pop_dens_friction.csv joined on
# data/raw/district_relation.csv NEW_DIST_ID_2); used only to shape synthetic populations.
DISTRICT_POP_60 = [
    629436, 2287502, 266877, 151584, 356768, 325633, 497716, 179697, 142384, 190749, 136070, 230535, 268506,
    243463, 83150, 136068, 137965, 324862, 270077, 155871, 234843, 166817, 171764, 129143, 80777, 324568,
    423211, 365685, 104577, 269311, 282476, 177553, 333276, 182316, 220158, 341560, 187648, 198861, 146635,
    65307, 157419, 86110, 113141, 118859, 101743, 123419, 82850, 141017, 102558, 82255, 97973, 65366, 87062,
    253838, 359959, 267520, 338602, 200036, 107597, 125905]
# Census frequencies of the economic statuses above (Teachers folded into Service Workers: the shipped
# 9-status matrix has no Teachers row).  notebooks/census/... Input.ipynb cell 5.
ECON_STATUS_FREQ = [3655252, 4603312, 1323535, 342493, 2235255 + 143152, 2407410, 126592, 43378, 177445]
