"""Mirror of ``pytorchltr.click_simulation`` (reference: pytorchltr/click_simulation/pbm.py)."""
from pytorchltr_b200.click_simulation.pbm import simulate_nearrandom  # noqa: F401
from pytorchltr_b200.click_simulation.pbm import simulate_pbm  # noqa: F401
from pytorchltr_b200.click_simulation.pbm import simulate_perfect  # noqa: F401
from pytorchltr_b200.click_simulation.pbm import simulate_position  # noqa: F401
