"""State-vector layout and normalisation constants.

Follows /root/reference/src/lua/config.lua:15-54 (0-based indices here, the Lua is 1-based).
"""
import math
import numpy as np

F32 = np.float32

VNC = 60.0            # velocity_normalize_constant            config.lua:16
ANC = math.pi         # angle_normalize_constant               config.lua:17
PNC = 800.0           # position_normalize_constant = max(cx,cy)*2   config.lua:48-51
MASSES = (1.0, 5.0, 25.0, 1e30)                               # config.lua:19
OBJECT_SIZES = (0.5, 1.0, 2.0)                                # config.lua:38
OID_BALL, OID_OBSTACLE, OID_BLOCK = 1, 2, 3                   # config.lua:41
BASE_SIZE = {"ball": 60.0, "obstacle": 80.0, "block": 60.0}   # config.lua:34
MAXWINSIZE = 60                                               # config.lua:28

# raw state indices (12 floats)   config.lua:22
R_PX, R_PY, R_VX, R_VY, R_A, R_AV, R_M, R_OID, R_OS, R_G, R_F, R_P = range(12)
RAW_DIM = 12

# one-hot state indices (22 floats)   config.lua:25
PX, PY, VX, VY, A, AV = range(6)
M0, M1 = 6, 10        # mass one-hot   [6,10)
OID0, OID1 = 10, 13   # oid one-hot    [10,13): ball, obstacle, block
OS0, OS1 = 13, 16     # size one-hot
G0, G1 = 16, 18
FR0, FR1 = 18, 20
P0, P1 = 20, 22
OBJ_DIM = 22
OSSI = 6              # object_state_start_index (args.ossi = si.m[1])   config.lua:52

NUM_PAST = 2          # main.lua:43
NUM_FUTURE = 1        # main.lua:44
INPUT_DIM = OBJ_DIM * NUM_PAST      # 44   main.lua:122
OUT_DIM = OBJ_DIM * NUM_FUTURE      # 22   main.lua:123
RNN_DIM = 50          # main.lua:37
LAYERS = 5            # main.lua:40
NBRHDSIZE = 3.5       # main.lua:39
MAX_CTX = 12          # data_sampler.lua:168, data_process.lua:59
BATCH_SIZE = 50       # main.lua:48
WINSIZE = 3           # num_past + num_future (main.lua:121)
