"""Host-side helpers mirroring ``qiskit_addon_obp.utils``."""
