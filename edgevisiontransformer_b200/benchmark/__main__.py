import sys

from .b200 import main

sys.exit(main())
