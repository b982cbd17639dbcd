"""Import-time stand-in for matplotlib.pyplot (only vehicle/dynamics.py's demo uses it)."""
