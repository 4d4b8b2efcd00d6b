from gymgo_b200.govars import *  # noqa: F401,F403
