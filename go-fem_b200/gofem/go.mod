module github.com/soypat/go-fem

go 1.18

require gonum.org/v1/gonum v0.11.0
