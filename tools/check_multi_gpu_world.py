"""tests/test_multi_gpu.py's two-rank check with any number of ranks (one per visible GPU): python tools/check_multi_gpu_world.py <world> <reduction>..."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gromacs-fep-gpu_b200", "python"), ROOT, os.path.join(ROOT, "tests")]


def main():
    import torch.multiprocessing as mp

    import test_multi_gpu as T

    world = int(sys.argv[1])
    for reduction in sys.argv[2:]:
        with tempfile.TemporaryDirectory() as d:
            result = os.path.join(d, "result.txt")
            mp.spawn(T._worker, args=(world, T._free_port(), ROOT, result, reduction), nprocs=world, join=True)
            print(f"world {world} {reduction}: {open(result).read()}", flush=True)


if __name__ == "__main__":
    main()
