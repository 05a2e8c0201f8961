"""``ase.constraints.FixConstraint`` placeholder.  TEST INFRASTRUCTURE."""


class FixConstraint:
    def index_shuffle(self, atoms, ind):
        raise NotImplementedError

    def adjust_positions(self, atoms, new):
        pass

    def adjust_momenta(self, atoms, momenta):
        pass
