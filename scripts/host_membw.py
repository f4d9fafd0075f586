"""Host memory copy bandwidth with T threads (numpy copies release the GIL): sizing input for a host-side
observation assembly in bd_step_host."""
import threading
import time

import numpy as np

for T in (1, 4, 8, 16):
    src = [np.random.rand(8 << 20).astype(np.float32) for _ in range(T)]      # 32 MB each
    dst = [np.empty_like(s) for s in src]

    def work(i):
        for _ in range(20):
            np.copyto(dst[i], src[i])

    th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    gb = T * 20 * src[0].nbytes / 1e9
    print(f"{T:2d} threads: {gb / dt:6.1f} GB/s copied (read + write traffic = 2x)")
