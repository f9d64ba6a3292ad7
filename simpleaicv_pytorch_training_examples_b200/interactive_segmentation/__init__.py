"""SAM image encoder of the interactive-segmentation hot path (SimpleAICV/interactive_segmentation)."""
