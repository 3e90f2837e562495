"""CPU oracle of the read-generation path: TEST INFRASTRUCTURE (checker only, never the thing measured or shipped)."""
