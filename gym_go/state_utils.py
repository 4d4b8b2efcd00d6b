from gymgo_b200.state_utils import *  # noqa: F401,F403
