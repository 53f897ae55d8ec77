"""Host-side state initialisation: factor distributions and sprite generators
(reference: moog/state_initialization/)."""
