"""oracle/ -- CPU checker (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  hyperpose_b200/ must never import it.
"""
from .binding import (OrcHuman, OrcPeak, OrcConn, build, load_oracle, load_ref, oracle_process,
                      ref_available, RefParser, resize_area_up, resize_area, gaussian17, gauss_kernel, area_up_tab, resize_linear_u8, pifpaf_ref_available, ref_pifpaf_process,
                      ppn_ref_available, ref_ppn_process)
from .ppn_oracle import ppn_process
