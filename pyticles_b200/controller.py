"""Controller base classes (controller.py:9-40) -- only what Box and the particle system
need; the one-body toy forces of the reference are outside the SPH hot path."""


class Controller(object):
    def __init__(self):
        self.groups = []

    def bind_particles(self, p):
        self.groups.append(p)

    def apply(self):
        print("Do nothing")
