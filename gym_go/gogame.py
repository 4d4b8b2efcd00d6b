from gymgo_b200.gogame import *  # noqa: F401,F403
from gymgo_b200.gogame import str  # noqa: F401,A004 - exported under this name by the reference
