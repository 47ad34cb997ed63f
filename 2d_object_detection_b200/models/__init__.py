"""Mirror of the reference's ``models`` package for the hot path (anchor helpers, ROIPooling)."""
