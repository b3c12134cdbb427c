"""Same module names as the reference's FeatureCraft-PyQt-App/utils package (drop-in imports)."""
