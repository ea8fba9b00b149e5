#!/usr/bin/env python
"""The reference's second example driver (README:89-101, ``<precision>/Examples/example_irl.F`` -- absent from the checkout): the
implicitly restarted xLANSVD_IRL on a matrix file, same options as example.py (``--kmax`` is the Krylov dimension ``dim``,
``--p`` the number of shifts per restart, ``--which S`` asks for the smallest triplets).

    python examples/example_irl.py tests/golden/illc1850.rra --k 10 --kmax 50 --p 40
"""
import sys

from example import main

if __name__ == "__main__":
    sys.exit(main(irl_default=True))
