"""Off-policy golden fixtures (filled in as the replay/PER/Q-learning kernels land).  TEST INFRASTRUCTURE."""


def main():
    pass
