"""oracle/ -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs).

ns3d_oracle.c + oracle.py   the CPU oracle: a literal, unfused C restatement of the reference's kernels and drivers
np_restatement.py           a second, independently written restatement (numpy slices)
jl_interp.py + jl_run.py    an interpreter for the Julia subset of the reference scripts: executes THEIR TEXT
                            (kernels, BC functions, parameter blocks, time loops, update_halo! call sites on several ranks)
jl_shim.py                  the same interpreter extended by structs / typed dispatch / Ref / ccall-through-ctypes:
                            executes julia/NS3DNative.jl and the re-pointed run scripts against the library's C ABI

Nothing under navierstokes3d_b200/ imports this package.
"""
