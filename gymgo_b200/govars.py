"""Channel constants of the 6xNxN state tensor - the values of the reference's gym_go/govars.py:1-11
(API constants; kept so `from gym_go import govars` keeps working)."""
ANYONE = None
NOONE = -1

BLACK = 0
WHITE = 1
TURN_CHNL = 2
INVD_CHNL = 3
PASS_CHNL = 4
DONE_CHNL = 5

NUM_CHNLS = 6
