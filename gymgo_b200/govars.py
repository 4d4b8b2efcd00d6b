"""Channel indices of the 6xNxN state tensor and the two player sentinels.

The VALUES are the reference's API constants (gym_go/govars.py:1-11) - user code indexes states with them
(`state[govars.INVD_CHNL]`), so they cannot differ; the packed record of the CUDA backend stores the same six
facts (gymgo_b200/csrc/gg_algo.cuh)."""

# stone planes, then the whole-plane facts: side to move, invalid-for-mover mask, previous-pass, game-over
BLACK, WHITE, TURN_CHNL, INVD_CHNL, PASS_CHNL, DONE_CHNL = range(6)
NUM_CHNLS = 6

# player sentinels
ANYONE, NOONE = None, -1
