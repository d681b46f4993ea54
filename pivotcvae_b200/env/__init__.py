"""Drop-in mirror of the reference's env/ package (user-response simulators)."""
