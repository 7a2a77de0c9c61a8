"""Input producers with the reference's module name (`data.load_data`): see load_data.py."""
